"""Generates tests/golden/shakespeare.txt.gz from the reference's benchmark corpus
(/root/reference/bench-data/shakespeare.txt, SURVEY.md §2 row 14: 5 465 394 bytes, sha256 8a304827...).
Run in the build container (the reference tree does not exist on the GPU box); the fixture travels with the repo.
BASELINE.json's configs name this file: configs[0] = the file x 100, configs[1..2] = the file x 10 000."""
import gzip
import hashlib
import os

SRC = "/root/reference/bench-data/shakespeare.txt"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shakespeare.txt.gz")
SHA256 = "8a304827e5ed421e8f7bfb0f66e8f87adf0130958cb88dcb187cbe06aea4be1f"

if __name__ == "__main__":
    data = open(SRC, "rb").read()
    assert len(data) == 5465394 and hashlib.sha256(data).hexdigest() == SHA256
    with open(DST, "wb") as f:
        with gzip.GzipFile(filename="", mode="wb", fileobj=f, compresslevel=9, mtime=0) as g:
            g.write(data)
    print(DST, os.path.getsize(DST))
