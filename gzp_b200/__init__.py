"""gzp_b200 — B200-native drop-in for gzp's per-block encode path.

The product is ``libgzpb.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/gzpb.h``).  This package is the host-side mirror of the reference's
public surface (``ZBuilder`` / ``ParCompressBuilder`` / ``ParCompress`` / the
format types / ``ZWriter::finish``; /root/reference/src/lib.rs:181-265,
src/par/compress.rs:33-233) so that tests read like the reference's own.
"""
from ._lib import BGZF, GZIP, MGZIP, RAWDEFLATE, SNAP, ZLIB, load  # noqa: F401
from .api import (BGZF_EOF, BgzfSyncWriter, MgzipSyncWriter, SyncZ, SyncZBuilder, BUFSIZE, DICT_SIZE, Bgzf, Compression, Context, Decoder, Gzip, GzpError, Mgzip,  # noqa: F401
                  NativeParCompress, NativeParDecompress, ParCompress, ParCompressBuilder, ParDecompress, ParDecompressBuilder, bgzf_index, bgzf_virtual_offset, compress_file, RawDeflate, Snap, ZBuilder, Zlib, adler32_combine, crc32_combine,
                  encode_capacity, encode_stream_multi, footer, header)
