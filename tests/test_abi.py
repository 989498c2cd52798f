"""The C-ABI library loads on a CPU-only box and exports every symbol include/iifb200.h declares;
without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from iifb200 import _abi as A

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "iifb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(iifb200_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_all_bound_in_python_abi():
    assert set(_declared_symbols()) == set(A.SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    lib = C.CDLL(A.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/iifb200.h but not exported"
    assert A.load_library().iifb200_version() == 200


def test_struct_sizes_match_header(built):
    """sizes computed by the C compiler for the oracle build must equal the ctypes mirrors"""
    import subprocess, tempfile, textwrap
    prog = textwrap.dedent("""
        #include <stdio.h>
        #include "iifb200.h"
        int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(iif_dist_desc), sizeof(iif_slot_desc),
          sizeof(iif_factor_desc), sizeof(iif_solver_params), sizeof(iif_conv_op), sizeof(iif_prop_op),
          sizeof(iif_product_op), sizeof(iif_sched_op), sizeof(iif_deconv_op), sizeof(iif_graph_desc),
          sizeof(iif_tree_desc), sizeof(iif_plan_opts)); return 0;}""")
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "s"), os.path.join(d, "s.c")])
        out = subprocess.check_output([os.path.join(d, "s")]).split()
    sizes = [int(x) for x in out]
    mirrors = [A.DistDesc, A.SlotDesc, A.FactorDesc, A.SolverParamsC, A.ConvOp, A.PropOp, A.ProductOp, A.SchedOp,
               A.DeconvOp, A.GraphDesc, A.TreeDesc, A.PlanOpts]
    assert sizes == [C.sizeof(m) for m in mirrors]


def test_no_gpu_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = A.load_library()
    ctx = C.c_void_p()
    st = lib.iifb200_init(0, C.byref(ctx))
    assert st != 0 and not ctx.value
    assert b"no CPU fallback" in lib.iifb200_last_error(None) or b"CUDA" in lib.iifb200_last_error(None)
    import parity_cases as PC
    P, xs, fs = PC.chain_problem(n=2, N=8)
    with pytest.raises(A.IIFB200Error):
        P.engine()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "incrementalinference.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "iif_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f
