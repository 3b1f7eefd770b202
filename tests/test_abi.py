"""CPU suite, part 2: the C-ABI library loads and exports every symbol include/gpr_c_api.h declares;
no compute entry point is called (there is no GPU here) except to check that it FAILS loudly."""
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _declared_symbols(path=None):
    text = open(path or os.path.join(ROOT, "include", "gpr_c_api.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gpr_[a-z_0-9]+)\s*\(", text)))


def test_header_and_library_agree(gpr):
    declared = _declared_symbols()
    assert sorted(gpr.C_ABI_SYMBOLS) == declared
    assert not [s for s in declared if "selftest" in s]       # probes live in the test-only header
    selftests = _declared_symbols(os.path.join(ROOT, "gaussian-object-modelling_b200", "csrc", "gpr_selftest.h"))
    assert sorted(gpr.SELFTEST_SYMBOLS) == selftests
    lib = gpr.lib()
    for s in declared + selftests:
        assert hasattr(lib, s), s
    out = subprocess.run(["nm", "-D", "--defined-only", gpr.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (gpr_[a-z_0-9]+)", out))
    assert exported == set(declared) | set(selftests)


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "gpr_c_api.h")).read()
    assert 'extern "C"' in text
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    assert "torch" not in code and "at::" not in code and "std::" not in code and "#include <stddef.h>" in code


def test_library_is_sm100a_only(gpr):
    out = subprocess.run(["cuobjdump", "--list-elf", gpr.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_fp64_tensor_and_async_copy_instructions_present(gpr):
    sass = subprocess.run(["cuobjdump", "-sass", gpr.LIB_PATH], capture_output=True, text=True).stdout
    assert sass.count("DMMA") > 500           # FP64 tensor pipe (mma.sync m8n8k4.f64 -> DMMA.8x8x4)
    assert "LDGSTS" in sass                   # cp.async staging of operand tiles


def test_product_never_touches_the_oracle():
    """The product path may not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "gaussian-object-modelling_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", "Makefile")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
                assert "oracle/" not in txt.replace("never imports the CPU oracle", ""), f
    for base, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            assert "oracle" not in open(os.path.join(base, f), errors="ignore").read(), f
    ldd = subprocess.run(["ldd", os.path.join(pkg, "libgpr_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in ldd


def test_fails_loudly_without_a_gpu(gpr):
    """No CPU fallback: on a box without a usable CUDA device the context cannot be created."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gpr.GPRegressionException) as e:
        gpr.Context()
    assert e.value.code == gpr.GPR_ERR_CUDA and "no CPU fallback" in str(e.value)


def test_header_is_plain_c99_and_links_from_c(gpr, tmp_path):
    """the boundary is usable from plain C.  The header must compile as strict C99 and a C program
    linked against the library must get the loud no-GPU failure (or a context, on a GPU box) — no C++ runtime types."""
    src = tmp_path / "abi_c99.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "gpr_c_api.h"\n'
        'int main(void) {\n'
        '    gpr_ctx* ctx = NULL;\n'
        '    gpr_kernel_t k; k.kind = 0; k.p0 = 4.2; k.p1 = 0.0;\n'
        '    int rc = gpr_ctx_create(NULL, 0, &ctx);\n'
        '    if (rc == GPR_OK) { printf("ctx devices=%d\\n", gpr_ctx_num_devices(ctx)); gpr_ctx_destroy(ctx); return 0; }\n'
        '    printf("rc=%d msg=%s kind=%d\\n", rc, gpr_last_error(), k.kind);\n'
        '    return rc == GPR_ERR_CUDA && strlen(gpr_last_error()) > 0 ? 0 : 1;\n'
        '}\n')
    pkg = os.path.join(ROOT, "gaussian-object-modelling_b200")
    exe = tmp_path / "abi_c99"
    cc = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                         "-o", str(exe), "-L", pkg, "-lgpr_b200", "-Wl,-rpath," + pkg], capture_output=True, text=True)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr


def test_python_mirror_argument_checks(gpr):
    """Same messages as the reference for the same conditions (gp_regressor.hpp:198,225,231,374,566,570);
    all raised before any GPU work."""
    reg = gpr.GPRegressor.__new__(gpr.GPRegressor)
    reg.set_cov_function("thin_plate", 2.0)
    with pytest.raises(gpr.GPRegressionException, match="Empty data pointer"):
        reg.create(None, None, None, None)
    with pytest.raises(gpr.GPRegressionException, match="All input data is empty!"):
        reg.create([], [], [], [])
    with pytest.raises(gpr.GPRegressionException, match="Empty Model pointer"):
        reg.evaluate(None, [0.0], [0.0], [0.0])
    with pytest.raises(gpr.GPRegressionException, match="Empty model pointer"):
        reg.update(None, [0.0], [0.0], [0.0], [0.0])
