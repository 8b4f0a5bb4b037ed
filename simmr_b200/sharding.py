"""Multi-GPU host logic: one process per GPU, reads sharded by global unit index.

The reference runs genomes sequentially in one thread (simulate.rs:135, main.rs:201-206);
its output order is the concatenation over genomes of their pairs / long reads.  Units are
independent, so rank k of W takes the contiguous range [k*N/W, (k+1)*N/W) of that order
with the genome set replicated on every GPU.  There is no data-path collective: the only
exchange is one all-gather of (bytes, records) per rank so every rank knows the file offset
of its shard and the single output file is byte-identical to the one-GPU file
(SURVEY.md §8e).  Works with any torch.distributed backend (nccl on the GPU box, gloo in
the CPU tests)."""
import os


def shard_range(total_units, rank, world):
    """Contiguous unit range of `rank`; ranges tile [0, total_units) in rank order."""
    if not 0 <= rank < world:
        raise ValueError("rank outside the world")
    return rank * total_units // world, (rank + 1) * total_units // world


def exchange_counts(nbytes, nrecords, dist=None):
    """All-gather of per-rank (bytes, records) -> (byte offset of this rank's shard,
    first record index of this rank, total bytes, total records)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return 0, 0, int(nbytes), int(nrecords)
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.tensor([int(nbytes), int(nrecords)], dtype=torch.int64, device=dev)
    every = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(every, mine)
    rank = dist.get_rank()
    table = [(int(t[0]), int(t[1])) for t in every]
    return (sum(b for b, _ in table[:rank]), sum(r for _, r in table[:rank]),
            sum(b for b, _ in table), sum(r for _, r in table))


class ShardWriter:
    """pwrite()s chunks of one rank's shard at its place in the shared output file."""

    def __init__(self, path, offset):
        self.fd = os.open(path, os.O_WRONLY | os.O_CREAT, 0o644)
        self.at = int(offset)

    def __call__(self, chunk, first_unit=0, n_units=0):
        view = memoryview(chunk)
        done = 0
        while done < len(view):
            done += os.pwrite(self.fd, view[done:], self.at + done)
        self.at += len(view)
        return 0

    def close(self):
        os.close(self.fd)
