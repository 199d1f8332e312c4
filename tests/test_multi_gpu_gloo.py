"""CPU, world_size 2 (gloo): the N > 1 host logic - clip sharding and the single all_gather of decoded frames."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch
sys.path.insert(0, os.environ["FYC_ROOT"])
from followyourclick_b200.distributed import init_from_env, shard_clips, gather_frames, gather_clip_results
d = init_from_env("gloo")
rank, world = d.get_rank(), d.get_world_size()
assert world == 2
assert shard_clips(5, rank, world) == ([0, 2, 4] if rank == 0 else [1, 3])
v = torch.full((1, 3, 2, 4, 4), float(rank + 1))
g = gather_frames(v)
assert g.shape == (2, 3, 2, 4, 4) and float(g[0].mean()) == 1.0 and float(g[1].mean()) == 2.0
mine = {i: torch.full((1, 3, 2, 4, 4), float(i)) for i in shard_clips(5, rank, world)}
allv = gather_clip_results(mine, 5)
assert len(allv) == 5 and all(float(allv[i].mean()) == float(i) for i in range(5))
d.barrier(); d.destroy_process_group()
print("ok", rank)
'''


def test_shard_and_gather_world_size_2(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   FYC_ROOT=ROOT, CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("ok" in o for o in outs), outs


def test_bench_reference_arm_ranks_other_than_zero_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], env=env,
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
