"""The reference arm of bench.py (`--impl reference`: the oracle port on the host threads) on a tiny workload: the JSON
contract of the line, and that ranks other than 0 exit quietly (under torchrun only rank 0 runs it)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ)
    env.update(env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "300", "--m", "600",
                           "--density", "0.05", "--steps", "4", "--warmup", "3"], capture_output=True, text=True, env=env,
                          cwd=ROOT, timeout=300)


def test_reference_arm_line():
    out = _run({"RANK": "0", "WORLD_SIZE": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "ADMM iterations/sec" and line["unit"] == "iter/s"
    assert line["higher_is_better"] is True and line["steps"] == 4 and line["warmup"] == 3 and line["value"] > 0
    assert abs(line["ms_per_step"] * line["value"] - 1e3) < 1e-6 * 1e3
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and cb["cores"] >= 1 and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["n"] == 300 and "workload" in line["config"]


def test_reference_arm_other_ranks_do_nothing():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_under_torchrun_prints_one_line():
    # the driver launches the reference arm like the engine arm: torchrun, one process per GPU; only rank 0 works
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "tests", "run_bench_small.py")], capture_output=True, text=True, env=env, cwd=ROOT,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["steps"] == 4
