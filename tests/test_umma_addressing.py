"""The tcgen05 prototypes under tools/microbench (round-2 groundwork, compiled but not yet run on a GPU) build their shared-
memory descriptors by hand; this keeps their offset arithmetic pinned against the documented canonical operand layout."""
import importlib.util
import os

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_prototype_descriptor_arithmetic():
    spec = importlib.util.spec_from_file_location("check_umma_addressing", os.path.join(REPO, "tools", "microbench", "check_umma_addressing.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    m.check_gru_layer(); m.check_gru_layer(units=96, k_in=352)
    m.check_conv_layer(); m.check_conv_layer(out=32, kc=192)
    m.check_gru_chain()
    m.check_tf32_refresh()


def test_weight_prebake_and_stage_sizing(tmp_path):
    """tools/microbench/umma_layout.h (the pre-bake that moves into weights.cpp with the tcgen05 codec): real layer shapes baked
    into 32 KB ring-stage chunks and read back through the planned tile descriptors, on the CPU"""
    import subprocess
    src = os.path.join(REPO, "tools", "microbench", "test_umma_layout.cpp")
    exe = str(tmp_path / "test_umma_layout")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, src], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "all layouts ok" in r.stdout, r.stdout
