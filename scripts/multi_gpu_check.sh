#!/bin/bash
# 2-GPU checks: the torchrun bench line, and the CLI's --gpus 2 output against --gpus 1 (must be byte-identical)
set -e
python - <<'PY'
import numpy as np
rng = np.random.default_rng(3)
for g in range(4):
    with open(f"/tmp/mg{g}.fna", "w") as f:
        for c in range(3):
            n = int(rng.integers(30000, 90000))
            s = "".join(np.array(list("ACGT"))[rng.integers(0, 4, n)])
            f.write(f">g{g}_c{c}\n")
            for i in range(0, n, 80):
                f.write(s[i:i + 80] + "\n")
PY
printf 'path\tuuid\n/tmp/mg0.fna\tu0\n/tmp/mg1.fna\tu1\n/tmp/mg2.fna\tu2\n/tmp/mg3.fna\tu3\n' > /tmp/mg.tsv   # fixed uuids: --genome alone draws random ones
G="--genome-file /tmp/mg.tsv"
./simmr_b200/simmr $G --output /tmp/one.fq --num-reads 2000001 --seed 11 --error-profile minimal-short --gpus 1
./simmr_b200/simmr $G --output /tmp/two.fq --num-reads 2000001 --seed 11 --error-profile minimal-short --gpus 2
cmp /tmp/one.fq /tmp/two.fq; cmp /tmp/one.fq.tsv /tmp/two.fq.tsv; echo "CLI --gpus 2 output identical to --gpus 1 ($(stat -c %s /tmp/one.fq) bytes)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0
