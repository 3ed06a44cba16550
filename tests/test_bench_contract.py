"""-m "not gpu": the parts of bench.py's contract that run without a GPU — the reference arm (the oracle port of the
reference's CPU path, `--impl reference`) prints one JSON line with the agreed keys, and under a multi-rank launch
only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip()


def test_reference_arm_line():
    text = _run()
    line = json.loads(text.splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "crystals/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["dtype"] == "f32" and line["data"] == "synthetic"
    # "reference" when oracle/_ref/cgat_reference.zip has been built (oracle/build_ref.py, where /root/reference exists)
    have_ref = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "cgat_reference.zip"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "cfg2_train" in line["config"]["workload"]


def test_reference_arm_other_ranks_stay_silent():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == ""
