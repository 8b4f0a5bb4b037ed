"""Synthetic genome sets of the sizes BASELINE.json's configs name: i.i.d. A,C,G,T, contig ids
g{genome:04}_c{contig:03}, uuids %016x of splitmix64(genome index).  Host-side input generation only — no part of
the simulation path."""
from concurrent.futures import ThreadPoolExecutor
import os

import numpy as np

from .api import GenomeSet

_SLAB = 1 << 26
_LETTER = np.frombuffer(bytes(b"ACGT"[i & 3] for i in range(256)), np.uint8)


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def random_bases(total, seed, threads=None):
    """`total` i.i.d. letters; slab k of 64 MB comes from its own SFC64 stream (seed, k), so the bytes do not depend on
    the thread count."""
    bases = np.empty(total, np.uint8)

    def slab(k):
        at = k * _SLAB
        m = min(_SLAB, total - at)
        words = np.random.Generator(np.random.SFC64([seed, k])).integers(0, 1 << 63, (m + 7) // 8, dtype=np.int64)
        np.take(_LETTER, words.view(np.uint8)[:m], out=bases[at:at + m])       # two bits of every byte of a draw

    with ThreadPoolExecutor(threads or min(16, os.cpu_count() or 1)) as ex:
        list(ex.map(slab, range((total + _SLAB - 1) // _SLAB)))
    return bases


def synthetic_genomes(n, lo, hi, contigs=(1, 1), seed=1, min_contig=50_000):
    """n genomes with sizes uniform in [lo, hi] and contigs[0]..contigs[1] contigs each, every contig >= min_contig."""
    rng = np.random.default_rng(seed)
    sizes = rng.integers(lo, hi + 1, n)
    total = int(sizes.sum())
    coff, csize, gcon, cids, cidoff, gids, gidoff = [0], [], [0], [], [0], [], [0]
    ncid = ngid = 0
    for g in range(n):
        nc = int(rng.integers(contigs[0], contigs[1] + 1))
        room = int(sizes[g]) - nc * min_contig
        if room < 0:
            raise ValueError("genome smaller than its contigs' minimum sizes")
        cuts = np.sort(rng.integers(0, room + 1, nc - 1)) if nc > 1 else np.zeros(0, np.int64)
        for c, ln in enumerate(np.diff(np.concatenate([[0], cuts, [room]])) + min_contig):
            coff.append(coff[-1] + int(ln))
            csize.append(int(ln))
            cid = f"g{g:04d}_c{c:03d}".encode()
            cids.append(cid)
            ncid += len(cid)
            cidoff.append(ncid)
        gcon.append(len(csize))
        gid = f"{splitmix64(g):016x}".encode()
        gids.append(gid)
        ngid += len(gid)
        gidoff.append(ngid)
    assert coff[-1] == total
    return GenomeSet(random_bases(total, seed), coff, csize, gcon, b"".join(cids), cidoff, b"".join(gids), gidoff,
                     genome_sizes=sizes.astype(np.uint64))


def lognormal_abundances(n, seed=3):
    """C3: abundances log-normal(sigma = 1) normalised to sum 1."""
    ab = np.random.default_rng(seed).lognormal(0, 1, n)
    return ab / ab.sum()
