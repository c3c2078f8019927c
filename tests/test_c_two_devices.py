"""GPU (2 devices): a plain C host -- no Python, no torch -- shards a fit over two GPUs through include/nk_b200.h alone:
per-device Gram passes, nk_allreduce_grams with the host's own NCCL communicators, column-sharded nk_solve_abc_part, finish.
tests/c/two_devices.c is compiled here with gcc against libnkb200.so, libcudart and the system libnccl.  Skipped on a one-GPU box."""
import pathlib
import shutil
import subprocess

import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
SRC = ROOT / "tests" / "c" / "two_devices.c"
CUDA = pathlib.Path("/usr/local/cuda")


def compile_program(out):
    lib = ROOT / "nys_koop_lqr_b200"
    cmd = ["gcc", "-std=gnu99", "-O1", str(SRC), "-o", str(out), f"-I{ROOT / 'include'}", f"-I{CUDA / 'include'}", f"-L{lib}", f"-L{CUDA / 'lib64'}",
           "-lnkb200", "-lcudart", "-lnccl", "-lm", f"-Wl,-rpath,{lib}", f"-Wl,-rpath,{CUDA / 'lib64'}"]
    return subprocess.run(cmd, capture_output=True, text=True)


def test_c_program_compiles_and_links(tmp_path):
    """CPU: the C host program builds against the public header and the shared library (no GPU needed to link)."""
    if shutil.which("gcc") is None or not (ROOT / "nys_koop_lqr_b200" / "libnkb200.so").exists():
        pytest.skip("gcc or libnkb200.so missing")
    r = compile_program(tmp_path / "two_devices")
    assert r.returncode == 0, r.stderr[-3000:]


@pytest.mark.gpu
def test_c_host_shards_a_fit_over_two_devices(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = tmp_path / "two_devices"
    r = compile_program(exe)
    assert r.returncode == 0, r.stderr[-3000:]
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600)
    print(run.stdout, run.stderr[-2000:])
    assert run.returncode == 0 and "OK" in run.stdout, (run.returncode, run.stdout[-2000:], run.stderr[-2000:])
