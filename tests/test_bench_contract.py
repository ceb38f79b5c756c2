"""bench.py's reference arm (`--impl reference`: the oracle port on the host cores) keeps the driver's JSON contract —
checked on CPU with a 64-byte message (N = M = 2^16) so the whole thing runs in seconds."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    env = dict(os.environ); env.pop("RANK", None); env.pop("WORLD_SIZE", None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--msg-len", "64"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    b = json.loads(line)
    assert b["impl"] == "reference" and b["metric"] == "sha256_r1cs_prove_field_ops_per_sec" and b["unit"] == "field-ops/s"
    assert b["higher_is_better"] is True and b["value"] > 0 and b["ms_per_step"] > 0 and b["steps"] == 1
    assert b["config"]["workload"] == "sha256_spartan_64B_zero_message" and b["config"]["N"] == 1 << 16
    cb = b["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "oracle" in cb["sample"] and cb["value"] == b["value"]
    assert b["e2e"] == {"value": b["value"], "unit": b["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_nonzero_ranks_do_no_work():
    env = dict(os.environ); env["RANK"] = "1"; env["WORLD_SIZE"] = "2"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--msg-len", "64"],
                         capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""
