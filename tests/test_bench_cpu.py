"""bench.py pieces that need no GPU: the reference (CPU) arm end to end at a reduced patch size, the clock-sample summary, the
staleness rule of the ncu traffic figure, the loud failure of the B200 arm without a device, and the bench-line contract on the
lines committed under profiles/."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _last_line(text):
    lines = [x for x in text.splitlines() if x.startswith("{")]
    assert lines, text[-400:]
    return json.loads(lines[-1])


def test_reference_arm_prints_one_contract_line_on_cpu():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cpu-size", "64", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-600:]
    d = _last_line(p.stdout)
    import bench

    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == (os.cpu_count() or 1) and cb["value"] == d["value"] and "64x64" in cb["sample"]
    assert "scaled x256" in cb["sample"]   # a reduced patch is declared as scaled; the driver's default run uses 1024 directly
    assert d["e2e"] == dict(value=d["value"], unit=bench.UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_b200_arm_fails_loudly_without_a_gpu():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600,
                       env=env, cwd=ROOT)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)


def test_clock_summary_prefers_nvidia_smi_and_falls_back_to_the_dense_series():
    import bench

    s = bench.ClockSampler.__new__(bench.ClockSampler)
    s.stop_flag = type("F", (), {"set": lambda self: None})()
    s.join = lambda timeout=None: None
    s.rows = [["1500", "1965", "900.1", "Not Active", "Not Active", "Not Active", "Active"]] * 3 + [["1400", "1965", "990", "Not Active", "Not Active", "Not Active", "Active"]]
    s.nvml_rows = [(1450.0, 1965.0, 0x4), (1300.0, 1965.0, 0x4 | 0x20)]
    out = s.summary()
    assert out["source"] == "nvidia-smi" and out["sm_mhz"] == 1500.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 4
    assert out["reasons"] == ["sw_power_cap", "sw_thermal_slowdown"] and out["nvml_samples"] == 2 and out["nvml_sm_min_mhz"] == 1300.0
    s.rows = s.rows[:1]
    out = s.summary()
    assert out["source"].startswith("nvml") and out["sm_mhz"] == 1375.0 and out["samples"] == 1
    s.rows, s.nvml_rows = [], []
    out = s.summary()
    assert out["sm_mhz"] is None and out["reasons"] == []


def test_traffic_figure_is_dropped_when_the_conv_source_changed(tmp_path, monkeypatch):
    import bench

    traffic, note = bench.read_traffic(16)
    src = os.path.join(ROOT, "kidney_diffusion_b200", "csrc", "kd_conv_gemm.cu")
    tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    if tj["conv_source_sha256"] == hashlib.sha256(open(src, "rb").read()).hexdigest():
        assert traffic == pytest.approx((tj["dram_bytes_read"] + tj["dram_bytes_write"]) * 16 / tj["batch"]) and "ncu" in note
    else:
        assert traffic is None and "stale" in note
    # a capture made from another version of the source must not be reported
    fake = tmp_path / "profiles"
    fake.mkdir()
    (fake / "r02_traffic.json").write_text(json.dumps(dict(tj, conv_source_sha256="0" * 64)))
    (tmp_path / "kidney_diffusion_b200" / "csrc").mkdir(parents=True)
    (tmp_path / "kidney_diffusion_b200" / "csrc" / "kd_conv_gemm.cu").write_bytes(open(src, "rb").read())
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.read_traffic(16)[0] is None


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_committed_bench_lines_obey_the_contract(n):
    import bench

    d = _last_line(open(os.path.join(ROOT, "profiles", f"r02_bench_n{n}.json")).read())
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "clocks", "cpu_baseline" if n == 1 else "identical_to_1gpu", "grid", "output"):
        assert k in d, k
    assert d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["n_gpus"] == n and d["scaling"] == "weak" and d["dtype"] == "f16"
    assert d["value"] > 0 and d["gpu_launches"] > 0 and "workload" in d["config"] and "model" not in d["config"]
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == bench.UNIT and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] <= d["value"] * 1.02
    assert not (set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}) and d["clocks"]["sm_mhz"] > 1000
    assert d["output"]["finite"] is True and d["grid"]["image_seconds_extrapolated"] > 0
    if n == 1:
        r = d["roofline"]
        assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.5 < r["frac"] < 1.0
        assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["parity"]["rel_l2_fp16_path_vs_fp32_path"] < d["parity"]["tolerance"]
    else:
        assert d["identical_to_1gpu"] is True
    # the reduced-step 16 384^2 image is the same image at every GPU count
    ref = _last_line(open(os.path.join(ROOT, "profiles", "r02_bench_n1.json")).read())["grid"]["reduced_run"]["checksum"]
    assert d["grid"]["reduced_run"]["checksum"] == ref
