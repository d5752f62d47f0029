"""Launched under torchrun by tests/test_bench_reference_arm_cpu.py: bench.py's reference arm on a tiny workload (torchrun's own
argument parser rejects bench.py's --n / --m as ambiguous prefixes of its options, so the sizes are set here)."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.argv = [os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--n", "300", "--m", "600", "--density", "0.05",
            "--steps", "4", "--warmup", "3"]
runpy.run_path(sys.argv[0], run_name="__main__")
