"""Repository-level contracts that can be checked without a GPU."""
import json
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_product_package_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under harmonypy_b200/ may import or reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "harmonypy_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(import oracle|from oracle)", text, flags=re.M):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_reference_arm_json_contract():
    """`bench.py --impl reference` prints one JSON line with the contract's keys (tiny sample)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                          "--cpu-sample", "3000", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in line, k
    assert line["impl"] == "reference" and line["vs_baseline"] is None and line["gpu_launches"] == 0
    # the staged unmodified reference (baseline/_ref, present wherever /root/reference is or was) or, without it, the port
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    if os.path.exists(os.path.join(ROOT, "baseline", "_ref", "harmonypy", "harmony.py")):
        assert line["cpu_baseline"]["kind"] == "reference"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0


def test_magic_division_is_exact_on_the_ranges_the_kernels_use():
    """hmy_magic / hmy_div (hmy_round_mma.cuh): floor(n * ceil(2^32 / d) / 2^32) == n // d for
    n < 2^24, 2 <= d <= 256 (d = 1 is special-cased in the kernel)."""
    rng = np.random.default_rng(0)
    n = np.concatenate([rng.integers(0, 1 << 24, size=200000, dtype=np.uint64),
                        np.array([0, 1, (1 << 24) - 1], dtype=np.uint64)])
    for d in list(range(2, 257)):
        m = np.uint64((0x100000000 + d - 1) // d)
        assert m < (1 << 32)
        q = (n * m) >> np.uint64(32)
        assert np.array_equal(q, n // np.uint64(d)), d


def test_design_and_integration_documents_exist_and_cite_the_reference():
    for name in ("DESIGN.md", "INTEGRATION.md"):
        text = open(os.path.join(ROOT, name)).read()
        assert "harmony.py:" in text and len(text) > 3000


def test_every_engine_option_and_counter_is_documented_in_the_header():
    """include/harmony_b200.h is the contract: each name hmy_set_option / hmy_counter / hmy_timer_ms accepts
    must appear in its comment block (force_fused_kernel is a codegen A/B switch, deliberately undocumented)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    api = open(os.path.join(root, "harmonypy_b200", "csrc", "hmy_api.cu")).read()
    header = open(os.path.join(root, "include", "harmony_b200.h")).read()

    def names_in(func):
        body = api[api.index(f'extern "C" {func}'):]
        body = body[:body.index("\n}\n")]
        return set(re.findall(r'n == "([a-z_0-9]+)"', body))

    opts = names_in("int hmy_set_option") - {"force_fused_kernel"}
    ctrs = names_in("int64_t hmy_counter")
    tmrs = names_in("double hmy_timer_ms")
    assert {"persistent", "mma", "tc5", "trace"} <= opts and {"launches", "grid", "tc5"} <= ctrs and "ms_round" in tmrs
    missing = sorted(n for n in opts | ctrs | tmrs if f'"{n}"' not in header)
    assert not missing, f"undocumented in include/harmony_b200.h: {missing}"
