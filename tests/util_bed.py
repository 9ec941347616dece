"""Test helpers for PLINK .bed images: a seeded generator (with missing genotypes) and a pure-Python restatement
of read_bed<char>() (/root/reference/src/read_bed.cpp:97-232), written independently of oracle/hb_oracle_ld.c
(dict-and-loop style, small sizes only)."""
import numpy as np

MAGIC = np.array([0x6C, 0x1B, 0x01], dtype=np.uint8)


def make_bed(nid, m, seed=0, p_missing=0.1, all_missing_cols=()):
    """Random SNP-major .bed image; returns (image bytes incl. magic, nid x m array of the 2-bit fields).
    About a third of the SNPs have no missing genotype at all."""
    rng = np.random.default_rng(seed)
    fields = np.empty((nid, m), dtype=np.uint8)
    for j in range(m):
        p = rng.uniform(0.05, 0.95)
        g = rng.binomial(2, p, size=nid)                  # genotype 0/1/2
        f = np.array([3, 2, 0], dtype=np.uint8)[g]        # -> field
        if j % 3 != 0:
            f[rng.random(nid) < p_missing] = 1
        if j in all_missing_cols:
            f[:] = 1
        fields[:, j] = f
    bps = (nid + 3) // 4
    body = np.zeros((m, bps), dtype=np.uint8)
    for i in range(nid):
        body[:, i // 4] |= fields[i, :] << (2 * (i % 4))
    # padding bits of the last byte: junk, the decoder must ignore them
    if nid % 4:
        junk = rng.integers(0, 4, size=m).astype(np.uint8)
        for x in range(nid % 4, 4):
            body[:, -1] |= junk << (2 * x)
    return np.concatenate([MAGIC, body.reshape(-1)]), fields


def py_read_bed(image, nid, m, impute, dominance, na=-128):
    code = {3: 0, 2: 1, 1: na, 0: (0 if dominance else 2)}
    bps = (nid + 3) // 4
    out = np.zeros((nid, m), dtype=np.int8, order="F")
    miss = np.zeros(m, dtype=np.uint8)
    for j in range(m):
        row = image[3 + j * bps: 3 + (j + 1) * bps]
        for i in range(nid):
            v = code[(int(row[i // 4]) >> (2 * (i % 4))) & 3]
            out[i, j] = v
            if v == na:
                miss[j] = 1
    if impute:
        for j in range(m):
            if not miss[j]:
                continue
            col = out[:, j]
            values = [0, 1] if dominance else [0, 1, 2]
            best, major = 0, 0
            for v in values:
                c = int((col == v).sum())
                if c > best:
                    best, major = c, v
            col[col == na] = major
    return out, miss
