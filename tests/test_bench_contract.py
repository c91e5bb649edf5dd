"""bench.py's reference arm and work accounting on the CPU (the GPU arm runs on the B200 box): the JSON line carries every key of
the measurement contract, and the FLOP formula reproduces SURVEY.md section 8d's per-clip figure."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0",
                          "--clip-seconds", "1"], capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_audio_seconds_per_second" and d["unit"] == "audio-s/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["unit"] == d["unit"] and cb["value"] == d["value"] and "x 1s clip" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith("{")]


def test_step_flop_formula_matches_survey_figure():
    sys.path.insert(0, ROOT)
    import bench
    from tiny_audio_b200.engine import PathDims
    d = PathDims(proj_hidden=2048)
    f = bench.path_flops(d, 30.0, 464, 65, 375)
    assert abs(f["reference"] / 1e12 - 3.49) < 0.01                      # 3.49 TFLOP per 30 s clip = 116 GFLOP per audio-second
    assert abs(f["reference"] / 30.0 / 1e9 - 116.2) < 0.2
    assert f["executed"] < f["reference"] and abs((f["reference"] - f["executed"]) - 4.0 * d.vocab * d.lm_dim * (464 - 65)) < 1.0
    g = bench.path_flops(d, 30.0, 464, 65, 375, train_lm=True)
    assert g["executed"] - f["executed"] > 2.0 * 440e6 * 464              # + the decoder's weight-gradient GEMMs


def test_clock_sampler_summarises_nvidia_smi_rows():
    sys.path.insert(0, ROOT)
    import bench

    class FakeProc:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0
    cs = bench.ClockSampler(0)
    cs.proc = FakeProc()
    cs.rows = [["0", "1530", "1965", "998.1", "0x4", "Not Active", "Not Active", "Not Active", "Active"],
               ["0", "1672", "1965", "1001.3", "0x4", "Not Active", "Not Active", "Not Active", "Active"],
               ["0", "1545", "1965", "1000.0", "0x0", "Not Active", "Not Active", "Not Active", "Not Active"],
               ["garbled"]]
    got = cs.stop()
    assert got == {"sm_mhz": 1545.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "samples": 3}
    cs2 = bench.ClockSampler(0)          # nvidia-smi could not be started: the line says so instead of inventing clocks
    assert cs2.stop()["reasons"] == ["nvidia-smi unavailable"]
