"""bench.py pieces that need no GPU: the stream windows and the reference arm's JSON line (the driver runs
`bench.py --impl reference` on the GPU box; its contract keys are held here on CPU)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_stream_windows_are_the_repeated_pattern():
    import bench
    from gzp_b200 import synth
    p = b"abcdefghij"
    assert bench.window(p, 7, 5) == b"hijab" and bench.window(p, 3, 4) == b"defg" and bench.window(p, 13, 27) == (p * 5)[3:30]
    c = synth.corpus()
    assert bench.window(c, len(c) - 3, 10) == c[-3:] + c[:7]
    assert synth.corpus_stream(10, len(c) - 3) == c[-3:] + c[:7]
    for name, cfg in bench.CONFIGS.items():
        assert cfg["blocks"] % cfg["inflight"] == 0 and cfg["block"] >= 32768, name      # whole device batches per step


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--blocks", "48"],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "bgzf_l6_compress_input_throughput" and line["unit"] == "GiB/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["dtype"] == "u8"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "GiB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "Bgzf" in line["config"]["workload"]
