"""world_size-2 gloo test of the N>1 plumbing (frame ownership, barrier, max/sum over
ranks) and of bench.py's reference arm contract under torchrun-style env."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
sys.path.insert(0, %r)
from human_body_proportion_estimation_b200 import dist_util
g = dist_util.Group(backend="gloo")
frames = dist_util.frames_of_rank(9, g.rank, g.world)
g.barrier()
mx = g.max_over_ranks([1.0 + g.rank, 5.0 - g.rank])
sm = g.sum_over_ranks([len(frames)])
print(json.dumps({"rank": g.rank, "frames": frames, "max": mx, "sum": sm}), flush=True)
g.close()
''' % ROOT


def test_two_rank_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = 29600 + os.getpid() % 300
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.PIPE, text=True))
    outs = []
    for p in procs:
        o, e = p.communicate(timeout=120)
        assert p.returncode == 0, e[-2000:]
        outs.append(json.loads(o.strip().splitlines()[-1]))
    outs.sort(key=lambda d: d["rank"])
    assert outs[0]["frames"] == [0, 2, 4, 6, 8] and outs[1]["frames"] == [1, 3, 5, 7]
    for d in outs:
        assert d["max"] == [2.0, 5.0]          # max over ranks, identical everywhere
        assert d["sum"] == [9.0]               # every frame owned exactly once


def test_reference_arm_json_contract():
    """bench.py --impl reference prints one valid JSON line (rank 0 only); other ranks exit 0 silently"""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["unit"] == "crops/s" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["config"]["workload"].startswith("configs[1]")
