"""bench.py contract checks that need no GPU: the reference arm (the real reference module on CPU) prints exactly one JSON line with the
keys the driver reads, and the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, cwd=ROOT, env=env)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "2", "--layers", "2")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["vs_baseline"] is None and "workload" in d["config"]
    # the real reference module when it is on the box (baseline/_ref or /root/reference), else the oracle port
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    sys.path.insert(0, ROOT)
    from oracle import ref_shim
    assert (d["cpu_baseline"]["kind"] == "reference") == ref_shim.reference_available()
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_product_arm_fails_loudly_without_a_gpu():
    r = _run("--steps", "1", "--warmup", "0", "--layers", "2", "--batch", "8")
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout) or "CUDA" in r.stderr
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
