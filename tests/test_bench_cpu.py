"""bench.py's output contract on the arm that runs without a GPU (`--impl reference`: the CPU restatement of the
reference step): stdout carries exactly ONE line, that line is the JSON record with the keys the driver reads --
whatever native libraries or warnings print goes to stderr."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_json_line_on_stdout():
    # a reduced scene keeps this to seconds; the contract (keys, one line) does not depend on the size
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--npoints", "3000", "--nqueries", "128"],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.splitlines()
    assert len(lines) == 1, lines
    rec = json.loads(lines[0])
    assert rec["impl"] == "reference" and rec["higher_is_better"] is True and rec["unit"] == "scenes/s"
    assert rec["value"] > 0 and rec["steps"] == 1 and rec["n_gpus"] == 1
    for key in ("metric", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in rec, key
    assert rec["cpu_baseline"]["kind"] == "port" and rec["cpu_baseline"]["cores"] >= 1
    assert rec["e2e"]["value"] == rec["value"] and rec["e2e"]["h2d_bytes_per_step"] == 0
    assert rec["config"]["workload"]
