"""CPU: host-side logic and the C ABI surface (no compute calls, no GPU needed)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from tests import _util as U


def _declared_c_symbols():
    """every function name the plain-C FFI header declares (after preprocessing)"""
    inc = os.path.join(U.ROOT, "include")
    src = subprocess.check_output(["gcc", "-E", "-P", "-x", "c", os.path.join(inc, "kblas_ffi.h")], text=True)
    names = set(re.findall(r"\b(kblas\w*)\s*\(", src))
    return sorted(names)


def test_library_loads_and_exports_every_declared_symbol(built):
    kb = U.kblas()
    lib = C.CDLL(kb.LIB_PATH)
    names = _declared_c_symbols()
    assert len(names) >= 50
    for fam in ("potrf", "trsm", "potrs", "posv"):
        for p in "SD":
            assert f"kblas{p}{fam}_batch" in names and f"kblas{p}{fam}_batch_strided" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_mangled_cpp_symbols_match_the_reference_abi(built):
    """code compiled against the reference headers links unchanged: same mangled names
    (reference include/kblas.h:54-108, kblas_batch.h wsquery + overloads, Xhelper_funcs.ch:48-55)"""
    kb = U.kblas()
    out = subprocess.check_output(["nm", "-D", "--defined-only", kb.LIB_PATH], text=True)
    ours = {l.split()[-1] for l in out.splitlines() if l.strip()}
    want = [
        U.mangle("kblasCreate", ["HH"]), U.mangle("kblasDestroy", ["HH"]),
        U.mangle("kblasAllocateWorkspace", ["H"]), U.mangle("kblasFreeWorkspace", ["H"]),
        U.mangle("kblasTimerTic", ["H"]), U.mangle("kblasTimerRecordEnd", ["H"]), U.mangle("kblasTimerToc", ["H"]),
        U.mangle("kblasCreateStreams", ["H", "i"]), U.mangle("kblasGetStream", ["H"]),
        U.mangle("kblasGetCublasHandle", ["H"]), U.mangle("kblasGetErrorString", ["i"]),
        "_Z14kblasSetStreamP11KBlasHandleP11CUstream_st",
        U.mangle("kblas_potrf_batch_wsquery", ["H", "i", "i"]),
        U.mangle("kblas_potrf_batch_strided_wsquery", ["H", "i", "i"]),
        U.mangle("kblas_trsm_batch_wsquery", ["H", "c", "i", "i", "i"]),
        U.mangle("kblas_trsm_batch_strided_wsquery", ["H", "c", "i", "i", "i"]),
        U.mangle("kblas_potrs_batch_wsquery", ["H", "i", "i", "i"]),
        U.mangle("kblas_potrs_batch_strided_wsquery", ["H", "i", "i", "i"]),
        U.mangle("kblas_posv_batch_wsquery", ["H", "c", "i", "i", "i"]),
        U.mangle("kblas_posv_batch_strided_wsquery", ["H", "c", "i", "i", "i"]),
        "_Z17kblas_potrf_batchP11KBlasHandleciPdiliPi", "_Z17kblas_potrf_batchP11KBlasHandleciPPdiiPi",
        "_Z17kblas_potrf_batchP11KBlasHandleciPfiliPi", "_Z17kblas_potrf_batchP11KBlasHandleciPPfiiPi",
        "_Z14Xset_pointer_1PPdPKdillP11CUstream_st", "_Z14Xset_pointer_1PPfPKfillP11CUstream_st",
        "_Z12iset_value_1PiilP11CUstream_st", "_Z8REG_SIZEi", "_Z16CLOSEST_REG_SIZEi",
    ]
    missing = [w for w in want if w not in ours]
    assert not missing, missing
    if U.have_ref():
        out = subprocess.check_output(["nm", "-D", "--defined-only", U.REF_SO], text=True)
        theirs = {l.split()[-1] for l in out.splitlines() if l.strip()}
        # every s/d symbol of the hot path the reference library exports: public calls, wsqueries, the internal offset
        # entry points, the pointer / value helpers -- non-uniform (MAGMA-only) trsm overloads included
        hot = [s for s in theirs
               if re.search(r"kblas_?(potrf|trsm|potrs|posv)_batch|kblas[SD](potrf|trsm|potrs|posv)_batch|"
                            r"X(potrf|potrs|posv)_batch_offset|Xtrsm_batch|iset_value_[1245]|Xset_pointer_[123]P", s)
               and "core" not in s
               and "cuComplex" not in s and "6float2" not in s and "7double2" not in s]
        absent = sorted(s for s in hot if s not in ours)
        assert not absent, absent


def test_return_code_constants_and_error_strings(built):
    kb = U.kblas()
    assert (kb.KBLAS_Success, kb.KBLAS_UnknownError, kb.KBLAS_NotImplemented, kb.KBLAS_InsufficientWorkspace) == (1, 0, -2, -6)
    assert kb.error_string(-6) == "Insufficient workspace supplied to function"   # kblas_common.cu:185
    assert kb.error_string(-2) == "Operation not implemented yet"
    assert kb.error_string(12345) == "unknown KBLAS error code"
    assert kb.roundup(33, 32) == 64 and kb.roundup(32, 32) == 32


def test_reg_size_rules(built):
    kb = U.kblas()
    # reference src/kblas_common.cu:241-255
    assert [kb.reg_size(n) for n in (0, 1, 2, 3, 16, 24, 32)] == [False, True, True, False, True, False, True]
    assert [kb.closest_reg_size(n) for n in (0, 1, 2, 3, 16, 17, 24, 32, 33, 100, 256)] == [0, 0, 1, 2, 8, 16, 16, 16, 32, 64, 128]


def _ref_potrf_ws(strided, n, batch):
    """restatement of workspace_queries.ch:111-122 + .cu:188-238 for the d_ptrs region"""
    def clos(x):
        r = 1
        while r < x:
            r <<= 1
        return r >> 1 if x > 0 else 0
    n1 = clos(n)
    need = 0
    if n1 > 16 and not strided:      # trsm(R, m = n-n1, n = n1) recursion -> offset GEMM pointer triples
        need = max(need, (batch > 1) * batch * 24)
    m = n - n1
    if m > 16:                        # syrk pointer triples, both modes
        depth, s = 0, 16
        while s < m:
            s <<= 1
            depth += 1
        need = max(need, (1 << (depth - 1)) * batch * 24)
    return need


@pytest.mark.parametrize("strided", [True, False])
@pytest.mark.parametrize("n", [8, 16, 24, 32, 33, 64, 100, 128, 256])
def test_potrf_wsquery_bytes(built, strided, n):
    kb = U.kblas()
    batch = 1000
    assert kb.wsquery_bytes("potrf", strided, 0, n, batch) == (0, 0, 0, _ref_potrf_ws(strided, n, batch))


def test_wsquery_headline_configs(built):
    kb = U.kblas()
    # strided potrf/trsm/potrs n<=32 -> 0 B; config 4 (ptr posv n=256, 64K) -> 6.3 MB d_ptrs
    assert kb.wsquery_bytes("potrf", True, 0, 32, 1 << 20) == (0, 0, 0, 0)
    assert kb.wsquery_bytes("trsm", True, 32, 32, 1 << 20, side="L") == (0, 0, 0, 0)
    assert kb.wsquery_bytes("potrs", True, 32, 32, 1 << 20) == (0, 0, 0, 0)
    assert kb.wsquery_bytes("posv", False, 16, 256, 1 << 16) == (0, 0, 0, 6291456)
    assert kb.wsquery_bytes("potrs", False, 16, 8, 100) == (0, 0, 0, 2400)      # offset GEMM even for n <= 16
    assert kb.wsquery_bytes("trsm", False, 32, 8, 100, side="R") == (0, 0, 0, 0)
    assert kb.wsquery_bytes("trsm", False, 32, 8, 100, side="L") == (0, 0, 0, 2400)
    assert kb.wsquery_bytes("trsm", False, 32, 8, 1, side="L") == (0, 0, 0, 0)    # (batchCount > 1) factor


def test_slab_partition():
    kb = U.kblas()
    slab = __import__("importlib").import_module("kblas-gpu_b200.slab")
    for batch, world in ((1 << 23, 8), (10, 4), (7, 8), (0, 2), (1000003, 3)):
        cover = []
        for r in range(world):
            b, e = slab.slab_range(batch, world, r)
            assert 0 <= b <= e <= batch
            cover += list(range(b, e)) if batch < 100 else []
            if r:
                assert b == slab.slab_range(batch, world, r - 1)[1]
        assert slab.slab_range(batch, world, world - 1)[1] == batch
        if batch < 100:
            assert cover == list(range(batch))
    assert slab.slab_offsets(1 << 23, 8, 3, 1024, 8) == (3 << 20, 1 << 20, (3 << 20) * 8192)
    assert kb is not None


def test_product_does_not_touch_the_oracle():
    """the shipped package must never import, link or fall back to oracle/ (tier rule 3)"""
    pkg = os.path.join(U.ROOT, "kblas-gpu_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) in ("build", "lib"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.lower() or f == "__init__.py" and "oracle" not in txt, (dirpath, f)
    out = subprocess.check_output(["ldd", U.kblas().LIB_PATH], text=True)
    assert "oracle" not in out and "openblas" not in out and "cublas" not in out


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU reference arm) runs without a GPU and prints ONE JSON line with the
    contract's keys; its e2e repeats its own value with zero copy bytes."""
    import json
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(U.ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=U.ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "matrices/s" and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "matrices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["value"] > 1e4


def test_potrf_smem_slot_plan_is_a_valid_colouring(built):
    """csrc/kernels/potrf_smem.cuh keeps the live 32 x 32 blocks of the factor in shared-memory slots assigned on the host.
    Replay the kernel's order of events and check that no two simultaneously live blocks share a slot, and that the slot
    count is the (nblk-J-1)(J+2) + (nblk-J-2) high-water mark (24 slots = 192 KiB for n = 256)."""
    kb = U.kblas()
    for nblk in range(1, 9):
        ns, slot = kb.potrf_smem_plan(nblk)
        live = {}

        def alloc(i, k):
            s = slot[i][k]
            assert s < ns and s not in live.values(), (nblk, i, k, s, live)
            live[(i, k)] = s

        for i in range(nblk):
            alloc(i, 0)
        for i in range(1, nblk):
            alloc(i, 1)
        high = len(live)
        for j in range(nblk):
            # step a reads (j, j) and (i, j) for i > j: all must be live
            assert all((i, j) in live for i in range(j, nblk))
            del live[(j, j)]
            for i in range(j + 2, nblk):
                alloc(i, j + 2)
            high = max(high, len(live))
            if j + 1 < nblk:
                # step b reads (i, k), i >= j+1, k <= j+1
                assert all((i, k) in live for i in range(j + 1, nblk) for k in range(j + 2))
                for k in range(j + 1):
                    del live[(j + 1, k)]
        assert not live
        assert ns == high
        want = max([(nblk - j - 1) * (j + 2) + max(0, nblk - j - 2) for j in range(nblk)] + [nblk + max(0, nblk - 1)])
        assert ns == want, (nblk, ns, want)
    assert kb.potrf_smem_plan(8)[0] == 24 and kb.potrf_smem_plan(4)[0] == 8 and kb.potrf_smem_plan(2)[0] == 3
