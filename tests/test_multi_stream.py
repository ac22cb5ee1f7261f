"""Multi-GPU path on CPU: world_size-2 `gloo` processes run bench.py's sharding/timing plumbing —
independent scan streams per rank, no data-path collective, barrier + max-over-ranks timing."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
import torch
import bench
from streammos_b200 import multi
world, rank, local = bench.dist_setup()
assert world == 2
# every rank owns whole sequences; together they cover all of them exactly once
mine = multi.sequences_of_rank(list(range(11)), rank, world)
import torch.distributed as dist
got = [None, None]
dist.all_gather_object(got, mine)
assert sorted(got[0] + got[1]) == list(range(11)) and not set(got[0]) & set(got[1])
# a rank's stream state is seeded by its rank: different streams, no shared state
assert multi.stream_seed(rank, 0) != multi.stream_seed(1 - rank, 0)
bench.barrier(world)
slow = bench.max_over_ranks(10.0 + 5.0 * rank, world, torch.device("cpu"))
assert slow == 15.0
total = multi.aggregate_scans_per_second(per_rank_steps=100, max_ms_total=slow * 100, world=world)
assert abs(total - 2 * 100 / (slow * 100 / 1e3)) < 1e-9
if rank == 0:
    print(json.dumps({"ok": True, "mine": mine, "slow": slow}))
dist.destroy_process_group()
''' % ROOT


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert '"ok": true' in r.stdout


def test_sequence_assignment_longest_first():
    from streammos_b200 import multi
    lengths = {"08": 4071, "11": 921, "12": 1061, "13": 3281, "14": 631, "15": 1901, "16": 1731, "17": 491,
               "18": 1801, "19": 4981, "20": 831, "21": 2721}
    world = 4
    parts = [multi.sequences_of_rank(list(lengths), r, world, lengths) for r in range(world)]
    assert sorted(sum(parts, [])) == sorted(lengths)
    loads = [sum(lengths[s] for s in p) for p in parts]
    assert max(loads) <= 1.25 * (sum(lengths.values()) / world)  # longest-first keeps ranks balanced
