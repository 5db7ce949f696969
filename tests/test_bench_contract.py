"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm
(`--impl reference`: the reference's own CPU path on the host cores) prints exactly one JSON line
on stdout with the keys the driver reads, whatever libraries write to file descriptor 1."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = out.stdout.splitlines()
    assert len(lines) == 1, out.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["metric"] == "displaced+shaded vertices/sec" and d["unit"] == "vertices/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout == ""


def test_both_arms_emit_the_same_config_object():
    """The driver compares the two arms' `config` objects: both come from make_config(world) and
    neither arm adds keys of its own afterwards."""
    import bench
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "config.update" not in src and src.count("make_config(world)") >= 2
    for world in (1, 2, 4, 8):
        c = bench.make_config(world)
        assert {"workload", "depth", "dim", "octaves", "gain", "faces", "quads", "vertices", "gpus", "quads_per_gpu",
                "vertices_per_gpu", "precision", "l2", "step"} <= set(c)
        assert c["quads"] == (16384 if world == 1 else 98304) and c["quads_per_gpu"] * world >= c["quads"]
