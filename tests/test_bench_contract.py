"""CPU: the parts of the bench.py contract that do not need a GPU -- the reference arm prints ONE JSON line with the
agreed keys (metric, unit, impl, cpu_baseline, e2e ...), non-zero ranks stay silent, and the product arm refuses to
run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                          timeout=600)


@pytest.fixture(scope="module")
def reference_line():
    from oracle import bindings
    if not os.path.exists(bindings.reference_path()) and not os.path.isdir("/root/reference/amcl3d/src"):
        pytest.skip("reference build unavailable")
    r = run_bench(["--impl", "reference", "--workload", "cfg2", "--steps", "1", "--warmup", "3", "--ref-particles", "40",
                   "--ref-procs", "2"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_line_has_the_contract_keys(reference_line):
    d = reference_line
    assert d["impl"] == "reference"
    assert d["metric"] == "particle_point_evals_per_s" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
    assert d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["value"] > 0 and d["ms_per_step"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] == 2 and cb["value"] == d["value"] and "sample" in cb
    assert cb["single_thread"]["cores"] == 1 and cb["single_thread"]["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    r = run_bench(["--impl", "reference", "--workload", "cfg2", "--steps", "1", "--warmup", "3"],
                  env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run_bench(["--workload", "cfg1", "--steps", "1", "--warmup", "3"])
    assert r.returncode != 0
    assert "no CPU fallback" in (r.stderr + r.stdout)


def test_default_workload_is_the_north_star_configuration():
    """No flags = configs[3]: 1 048 576 particles x 32 768 points on map L, strong scaling."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'ap.add_argument("--workload", default="cfg4"' in src
    assert '"scaling": "strong" if strong else "weak"' in src and "strong = not args.weak" in src
