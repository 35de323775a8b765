"""bench.py contract pieces that need no GPU: the reference arm's JSON line, the synthetic-weights helper, FLOP accounting."""
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_exactly_one_json_line():
    """`bench.py --impl reference` (the driver's reference arm): ONE line on stdout whatever libraries print, with the keys the
    contract names; it runs the oracle port of the reference's per-frame CPU path on every core of the affinity mask."""
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--height", "64", "--width", "96", "--gallery-rows", "500"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["metric"] == json.loads((ROOT / "BASELINE.json").read_text())["metric"]
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_synthetic_weights_equal_the_oracle_helper():
    """The CUDA arm of bench.py takes its weights from the package (it must not import oracle/); same seed -> same state dict."""
    from oracle import common
    from vision_sam3_yolo_lameless_b200.synthetic import random_init_vit

    a, b = random_init_vit("vitb16", seed=0, layers=1).state_dict(), common.hf_model(layers=1).state_dict()
    assert a.keys() == b.keys()
    assert all(torch.equal(a[k], b[k]) for k in a)


def test_flop_accounting_matches_survey():
    from vision_sam3_yolo_lameless_b200.engine import VitConfig
    from vision_sam3_yolo_lameless_b200.synthetic import VIT_SHAPES, vit_flops_per_frame

    assert abs(vit_flops_per_frame("vitb16", 201, 196) / 1e9 - 35.864) < 1e-3
    assert abs(vit_flops_per_frame("vitl16", 201, 196) / 1e9 - 125.68) < 1e-2
    assert abs(vit_flops_per_frame("vitb16", 1029, 1024) / 1e9 - 215.04) < 1e-2
    for name, (hidden, mlp, layers, heads) in VIT_SHAPES.items():
        cfg = VitConfig(hidden=hidden, layers=layers, heads=heads, mlp=mlp, patch=16, registers=4, rope_theta=100.0, ln_eps=1e-5)
        assert cfg.flops_per_frame(14, 14) == vit_flops_per_frame(name, 201, 196)


def test_bench_cuda_arm_does_not_import_oracle():
    """Only the cpu_baseline / reference legs of bench.py may touch oracle/."""
    src = (ROOT / "bench.py").read_text()
    arm = src[src.index("def run_b200("):src.index("_JSON_FD = None")]
    assert "from oracle" not in arm and "import oracle" not in arm


def test_bind_host_to_gpu_is_best_effort():
    from vision_sam3_yolo_lameless_b200.sharded import bind_host_to_gpu

    if not torch.cuda.is_available():
        assert bind_host_to_gpu(0) is None          # no device, no NVML: never raises
