"""bench.py contract on a CPU-only box: the reference arm (the reference's algorithm on the host cores — the oracle
port, the one other place bench.py may execute oracle/) prints the JSON line the driver parses; the product arm has no CPU
fallback and fails loudly without a GPU."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_line():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "rays/sec" and line["unit"] == "rays/s"
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    assert line["higher_is_better"] is True and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert line["value"] > 0 and line["ms_per_step"] > 0
    assert "configs[1]" in line["config"]["workload"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "rays" in cb["sample"]
    e = line["e2e"]
    assert e["value"] == line["value"] and e["unit"] == line["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful on a box without a GPU")
def test_product_arm_has_no_cpu_fallback():
    p = _run(["--steps", "1", "--warmup", "0", "--no-extras", "--no-cpu-baseline"], timeout=300)
    out = p.stdout.strip().splitlines()
    assert p.returncode != 0
    assert not any(l.startswith("{") and '"value"' in l for l in out)
