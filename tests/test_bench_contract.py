"""bench.py contract checks that need no GPU: the reference arm (the reference's CPU path, oracle/ as the timed
restatement) prints exactly ONE JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "audio-s/s"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C2")
    # both arms print the IDENTICAL config dict (the driver compares them) and the CPU model string rides along
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.config_dict(1)
    assert d["cpu_baseline"]["cpu"] and "sample_per_step" in d["cpu_baseline"]


def test_every_leg_runs_a_rank_independent_number_of_steps():
    """one collective per step: a step count derived from a rank-local timing deadlocks the ranks (it did once).  The native arm
    may size a leg from a timing only after a MAX all-reduce of that timing."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    i = src.index("# ---- sustained")
    leg = src[i:src.index("# ---- e2e through the host-buffer API")]
    assert "dist.all_reduce(tper, op=dist.ReduceOp.MAX)" in leg and leg.index("all_reduce(tper") < leg.index("ns = int(")


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
