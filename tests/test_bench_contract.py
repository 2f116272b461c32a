"""bench.py's contract, checked without a GPU: the reference arm's JSON line (the one arm that runs on the host), the
workload string both arms print, and the helpers around the timed regions."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    sys.path.insert(0, ROOT)
    import bench as b
    return b


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference`: the oracle port on the host threads, >= 10 s of CPU work whatever --steps says, same
    metric / unit / workload string as the GPU arm, e2e object that repeats the line's own value, no GPU needed."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-400:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert "DartHopper-v1, 4096 worlds/GPU" in d["config"]["workload"] and "10." in d["config"]["timed"]


def test_both_arms_name_the_same_workload(bench):
    from dart_env_b200.tasks import SPECS
    for cid, cfg in bench.CONFIGS.items():
        s = bench.workload_string(cfg["env"], cfg["worlds"], SPECS[cfg["env"]].task.frame_skip, cfg["lcp"], cfg["pgs_iters"])
        assert cfg["env"] in s and str(cfg["worlds"]) in s and ("pgs(30)" in s) == (cfg["lcp"] == "pgs")
    assert bench.workload_string("DartHopper-v1", 4096, 4, "exact", 30) == \
        "DartHopper-v1, 4096 worlds/GPU, fp32 engine, frame_skip 4, random actions U(-1,1), auto-reset, lcp=exact"


def test_configs_follow_baseline_json(bench):
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    text = json.dumps(base)
    for cid, cfg in bench.CONFIGS.items():
        assert cfg["env"].replace("-v1", "") in text.replace("-v1", "") or cfg["env"] in text, cfg["env"]
    assert bench.CONFIGS[2]["worlds"] == 4096 and bench.CONFIGS[3]["worlds"] == 16384 and bench.CONFIGS[4]["worlds"] == 16384


def test_affinity_helper_degrades_without_a_gpu(bench, monkeypatch):
    class _NoTorch:      # the helper must never take the bench down: no NVML / no GPU -> "unavailable", restore() is a no-op
        class cuda:
            @staticmethod
            def get_device_properties(i):
                raise RuntimeError("no CUDA")
    a = bench.GpuCpuAffinity(_NoTorch, 0)
    assert a.note.startswith("unavailable") or a.note.startswith("nvml")
    before = os.sched_getaffinity(0)
    a.restore()
    assert os.sched_getaffinity(0) == before
    monkeypatch.setenv("BENCH_CPU_AFFINITY", "0")
    assert bench.GpuCpuAffinity(_NoTorch, 0).note == "off"


def test_algorithmic_bytes_follow_survey_8d(bench):
    # 4 * (2 nd + nact + 2 nd + nobs + 1) + 1 bytes per env step; Hopper nd 6, nact 3, nobs 11 -> 157
    assert bench.algo_bytes(6, 3, 11) == 157
    assert bench.algo_bytes(9, 6, 17) == 241


def test_reference_arm_under_torchrun_prints_once():
    """N > 1: the driver launches the reference arm like the GPU arm (torchrun, one process per GPU); rank 0 alone runs and
    prints the line, the other ranks exit 0 without work."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=400)
    assert out.returncode == 0, out.stderr[-600:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["value"] > 0
