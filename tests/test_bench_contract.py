"""bench.py contract pieces that run without a GPU: the reference arm's JSON line and the CPU-path helper."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "reactor-solves/sec" and line["unit"] == "solves/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None and line["dtype"] == "f64"
    # kind "reference": the unmodified PyREMOT.rmtExe (from /root/reference, or its copy oracle/_ref on the GPU box); "port" only
    # where neither exists
    from oracle import ref_driver
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_driver.available() else "port")
    assert line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_copy_recipe_is_byte_identical(tmp_path):
    """oracle/build_ref.py places an unmodified copy of the reference package under oracle/_ref (build container only)."""
    import filecmp
    from oracle import build_ref
    if not os.path.isdir(build_ref.SRC):
        return
    assert build_ref.build(verbose=False)
    cmp = filecmp.dircmp(build_ref.SRC, os.path.join(build_ref.DST, "PyREMOT"), ignore=["__pycache__"])

    def walk(c):
        assert not c.diff_files and not c.left_only and not c.right_only, (c.left, c.diff_files, c.left_only, c.right_only)
        for sub in c.subdirs.values():
            walk(sub)
    walk(cmp)


def test_non_zero_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0 and "CUDA device" in out.stderr
