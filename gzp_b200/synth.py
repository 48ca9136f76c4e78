"""Deterministic synthetic corpora for tests and bench.py (no reference data travels
to the GPU box).  Shapes follow BASELINE.json / SURVEY.md §8d:

* ``text(n)``      — English-play-shaped ASCII text standing in for
                     bench-data/shakespeare.txt (≈4.6 bit/B order-0 entropy, DEFLATE
                     level-6 ratio ≈0.39): Zipfian vocabulary, recurring phrases,
                     speaker headings, short lines.
* ``low_entropy(n)`` — run-structured binary (config 4).
* ``fastq(n)``     — FASTQ-shaped records (config 5).
"""
import numpy as np

TEXT_PERIOD = 5465394  # bytes of the reference corpus (bench-data/shakespeare.txt)


def _vocab(rng, nwords):
    cons = ["b", "c", "d", "f", "g", "h", "l", "m", "n", "p", "r", "s", "t", "v", "w", "th", "st", "sh", "ch", "wh", "pr", "tr"]
    vows = ["a", "e", "i", "o", "u", "ou", "ea", "ee", "ai", "oo"]
    ends = ["", "", "", "s", "e", "ed", "er", "ing", "ly", "est", "th", "d", "n", "t"]
    common = ["the", "and", "I", "to", "of", "a", "you", "my", "that", "in", "is", "not", "with", "it", "me", "for",
              "be", "his", "your", "this", "he", "but", "have", "as", "thou", "him", "so", "will", "what", "thy",
              "all", "her", "no", "by", "do", "shall", "if", "are", "we", "thee", "on", "our", "lord", "king",
              "good", "now", "sir", "from", "come", "at", "they", "well", "she", "or", "let", "would", "more",
              "was", "here", "then", "love", "how", "am", "man", "their", "when", "there", "hath", "them", "one"]
    words = list(common)
    seen = set(words)
    while len(words) < nwords:
        k = 1 + int(rng.integers(0, 3))
        w = "".join(cons[int(rng.integers(len(cons)))] + vows[int(rng.integers(len(vows)))] for _ in range(k))
        w += ends[int(rng.integers(len(ends)))]
        if w not in seen:
            seen.add(w)
            words.append(w)
    return words


_CACHE = {}


def text(nbytes, seed=0x5EED0001):
    """`nbytes` of synthetic play text; deterministic in (nbytes, seed)."""
    key = (nbytes, seed)
    if key in _CACHE:
        return _CACHE[key]
    rng = np.random.default_rng(seed)
    words = _vocab(rng, 24000)
    ranks = np.arange(1, len(words) + 1, dtype=np.float64)
    p = 1.0 / ranks ** 1.07
    p /= p.sum()
    # recurring phrases (2-5 words), themselves Zipf-distributed
    nphr = 6000
    phr_len = rng.integers(2, 6, size=nphr)
    phr_words = rng.choice(len(words), size=int(phr_len.sum()), p=p)
    phr_off = np.concatenate([[0], np.cumsum(phr_len)])
    pp = 1.0 / np.arange(1, nphr + 1, dtype=np.float64) ** 0.9
    pp /= pp.sum()
    names = [w.upper() for w in words[200:260]]
    out = bytearray()
    target = nbytes + 4096
    line_len = 0
    while len(out) < target:
        # a speech: heading + a few lines
        out += b"\n" + names[int(rng.integers(len(names)))].encode() + b":\n"
        nw = int(rng.integers(8, 70))
        wi = rng.choice(len(words), size=nw, p=p)
        usephr = rng.random(nw) < 0.22
        phr_pick = rng.choice(nphr, size=nw, p=pp)
        cap = True
        line_len = 0
        for j in range(nw):
            if usephr[j]:
                a, b = phr_off[phr_pick[j]], phr_off[phr_pick[j] + 1]
                toks = [words[int(x)] for x in phr_words[a:b]]
            else:
                toks = [words[int(wi[j])]]
            for t in toks:
                if cap:
                    t = t[0].upper() + t[1:]
                    cap = False
                out += t.encode()
                line_len += len(t) + 1
                r = rng.random()
                if r < 0.07:
                    out += b","
                elif r < 0.11:
                    out += b"."
                    cap = True
                elif r < 0.125:
                    out += b"?" if r < 0.118 else b"!"
                    cap = True
                elif r < 0.135:
                    out += b";"
                if line_len > 38 + (j % 9):
                    out += b"\n"
                    line_len = 0
                else:
                    out += b" "
        out += b".\n"
    res = bytes(out[:nbytes])
    _CACHE[key] = res
    return res


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def low_entropy(nbytes, seed=0x5EED0004):
    """Run-structured low-entropy binary (≈1 bit/byte): value 0x00 p=.5, 0xFF p=.2, one of
    16 fixed bytes p=.3; run length 1..64."""
    rng = np.random.default_rng(seed)
    nruns = nbytes // 16 + 64
    r = rng.random(nruns)
    fixed = np.array([0x01, 0x02, 0x10, 0x20, 0x3C, 0x40, 0x55, 0x7F, 0x80, 0xAA, 0xC3, 0xE0, 0xF0, 0xFE, 0x33, 0x0F], dtype=np.uint8)
    vals = np.where(r < 0.5, 0, np.where(r < 0.7, 0xFF, fixed[rng.integers(0, 16, size=nruns)])).astype(np.uint8)
    lens = 1 + rng.integers(0, 64, size=nruns)
    out = np.repeat(vals, lens)
    while out.size < nbytes:
        out = np.concatenate([out, out])
    return out[:nbytes].tobytes()


def fastq(nbytes, seed=0x5EED0005):
    """FASTQ-shaped records: header, 150 bases, '+', 150 position-dependent qualities."""
    rng = np.random.default_rng(seed)
    nrec = nbytes // 320 + 2
    out = bytearray()
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)
    pos = np.arange(150)
    for i in range(nrec):
        out += f"@SYN.{i} {i}/1\n".encode()
        b = bases[rng.integers(0, 4, size=150)].copy()
        b[rng.random(150) < 0.001] = ord("N")
        out += b.tobytes() + b"\n+\n"
        q = np.clip(38 - pos // 12 - rng.geometric(0.35, size=150), 2, 40) + 33
        out += q.astype(np.uint8).tobytes() + b"\n"
        if len(out) >= nbytes:
            break
    return bytes(out[:nbytes])


def text_stream(nbytes, seed=0x5EED0001):
    """`nbytes` of the TEXT_PERIOD-byte synthetic corpus repeated, exactly as the reference
    bench concatenates shakespeare.txt xN (benches/bench.rs, README.md:166-167)."""
    import os
    cache = f"/tmp/gzpb_text_{TEXT_PERIOD}_{seed:x}.bin"
    if os.path.exists(cache) and os.path.getsize(cache) == TEXT_PERIOD:
        base = open(cache, "rb").read()
    else:
        base = text(TEXT_PERIOD, seed)
        try:
            tmp = f"{cache}.{os.getpid()}.tmp"          # several ranks may build the cache at once
            with open(tmp, "wb") as f:
                f.write(base)
            os.replace(tmp, cache)
        except OSError:
            pass
    reps = nbytes // TEXT_PERIOD + 1
    return (base * reps)[:nbytes]


CORPUS_BYTES = 5465394
CORPUS_SHA256 = "8a304827e5ed421e8f7bfb0f66e8f87adf0130958cb88dcb187cbe06aea4be1f"
_corpus = None


def corpus():
    """The reference's benchmark corpus itself — bench-data/shakespeare.txt (5 465 394 bytes), the file every
    BASELINE.json config names — from the committed fixture tests/golden/shakespeare.txt.gz
    (tests/golden/make_corpus.py made it; the reference tree does not exist on the GPU box)."""
    global _corpus
    if _corpus is None:
        import gzip
        import hashlib
        import os
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "shakespeare.txt.gz")
        data = gzip.open(path, "rb").read()
        if len(data) != CORPUS_BYTES or hashlib.sha256(data).hexdigest() != CORPUS_SHA256:
            raise RuntimeError("tests/golden/shakespeare.txt.gz does not hold the reference corpus")
        _corpus = data
    return _corpus


def corpus_stream(nbytes, start=0):
    """`nbytes` of shakespeare.txt repeated end to end, from stream offset `start` — BASELINE.json's
    "shakespeare x N" streams (benches/bench.rs, README.md:166-167), any window of them."""
    base = corpus()
    start %= CORPUS_BYTES
    reps = (start + nbytes) // CORPUS_BYTES + 1
    return (base * reps)[start:start + nbytes]
