"""CPU-side tests of the host layer: the C-ABI library loads and exports every
symbol the header declares, fails loudly without a GPU (no CPU fallback), and
the host-only logic (tau tail, slab partition, solver-name mapping, workload
generator) behaves like the reference."""
import math
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "openimpala_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    # oi_* plus the reference's own bind(c) names (tortuosity_fillmtx, tortuosity_remspot)
    return sorted(set(re.findall(r"\b((?:oi|tortuosity)_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built_lib):
    from openimpala_b200 import capi
    declared = _header_symbols()
    assert declared, "no declarations parsed"
    assert sorted(capi.EXPORTED_SYMBOLS) == declared
    for name in declared:
        assert getattr(built_lib, name) is not None
    assert built_lib.oi_version() == 100


def test_library_is_self_contained(built_lib):
    """cudart is linked statically and NCCL is loaded lazily: the .so must not
    need libcudart / libnccl at load time."""
    import subprocess
    from openimpala_b200 import capi
    out = subprocess.run(["ldd", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcudart" not in out and "libnccl" not in out and "not found" not in out


def test_sass_is_sm100a(built_lib):
    import shutil
    import subprocess
    from openimpala_b200 import capi
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device every compute entry point must fail loudly."""
    from openimpala_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.OiError) as e:
        capi.count_phase(np.zeros(8, dtype=np.uint8), 0)
    assert "no CPU fallback" in str(e.value)
    with pytest.raises(capi.OiError):
        capi.Solver((4, 4, 4), 0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "openimpala_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".H")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                # build.py may *compile* the checker (build_oracle); nothing may load or call it
                bad = re.search(r"import\s+oracle|from\s+oracle|liboi_oracle|oi_numpy|oi_c\b|oo_[a-z_]+\(", txt)
                assert not bad, f"{f} uses the oracle: {bad.group(0)}"


def test_tau_tail_conventions():
    from openimpala_b200.tortuosity import tau_from_fluxes
    # uniform 8-cell column: flux = A * (vhi-vlo)/(N-1) -> tau = (N-1)/N
    n = 8
    tau, deff, ok = tau_from_fluxes(-64.0 / 7.0, -64.0 / 7.0, 1.0, 8.0, 64.0, 0.0, 1.0)
    assert ok and abs(tau - (n - 1) / n) < 1e-15 and abs(deff - n / (n - 1)) < 1e-15
    assert math.isnan(tau_from_fluxes(-1.0, -1.00001, 0.5, 8.0, 64.0, 0.0, 1.0)[0])      # gate 1e-6
    assert not math.isnan(tau_from_fluxes(-1.0, -1.0000005, 0.5, 8.0, 64.0, 0.0, 1.0)[0])
    assert math.isinf(tau_from_fluxes(0.0, 0.0, 0.5, 8.0, 64.0, 0.0, 1.0)[0])
    assert math.isnan(tau_from_fluxes(0.0, 0.0, 0.0, 8.0, 64.0, 0.0, 1.0)[0])
    assert math.isinf(tau_from_fluxes(-1e-20, -1e-20, 0.5, 8.0, 64.0, 0.3, 0.3)[0])


def test_tau_tail_matches_oracle():
    from openimpala_b200.tortuosity import tau_from_fluxes
    from oracle import oi_numpy as o
    rng = np.random.default_rng(0)
    for _ in range(200):
        fin = -abs(rng.normal()) * 10
        fout = fin * (1 + rng.choice([0, 1e-7, 3e-6]))
        avf = rng.choice([0.0, 0.3, 1.0])
        shape = (5, 6, 7)
        d = int(rng.integers(0, 3))
        vlo, vhi = rng.choice([(-1.0, 1.0), (0.0, 1.0), (0.5, 0.5)])
        ext = (7.0, 6.0, 5.0)
        length = ext[d]
        area = ext[1] * ext[2] if d == 0 else (ext[0] * ext[2] if d == 1 else ext[0] * ext[1])
        a = tau_from_fluxes(fin, fout, avf, length, area, vlo, vhi)[0]
        b = o.tau_from_fluxes(fin, fout, avf, shape, d, vlo, vhi)[0]
        assert (math.isnan(a) and math.isnan(b)) or a == b


def test_solver_type_strings():
    from openimpala_b200.tortuosity import SolverType, string_to_solver_type
    assert [t.name for t in SolverType] == ["Jacobi", "GMRES", "FlexGMRES", "PCG", "BiCGSTAB", "SMG", "PFMG"]
    assert string_to_solver_type("flexgmres") is SolverType.FlexGMRES
    assert string_to_solver_type("PFMG") is SolverType.PFMG
    with pytest.raises(ValueError):
        string_to_solver_type("cg")


def test_slab_partition():
    from openimpala_b200 import capi
    for nz, n in [(1024, 8), (1024, 3), (100, 2), (2048, 8), (64, 4), (37, 2)]:
        parts = capi.slab_partition(nz, n)
        assert len(parts) == n and parts[0][0] == 0
        assert sum(p[1] for p in parts) == nz
        for (z0, m), (z1, _) in zip(parts, parts[1:]):
            assert z0 + m == z1 and z1 % 2 == 0 and m >= 2
    assert capi.slab_partition(50, 1) == [(0, 50)]


def test_sphere_packing_slabs_agree():
    from openimpala_b200 import synth
    full = synth.sphere_packing_slab((48, 40, 56), seed=3, radius=5)
    parts = [synth.sphere_packing_slab((48, 40, 56), seed=3, radius=5, z_begin=z, nz_local=16) for z in (0, 16, 32)]
    assert np.array_equal(full, np.concatenate(parts))
    assert 0.3 < full.mean() < 0.55
    assert synth.describe(full)["sha256"] == synth.describe(np.concatenate(parts))["sha256"]


def test_ctypes_struct_layout_matches_the_header(tmp_path):
    """oi_params / oi_solve_info as the C compiler lays them out == the ctypes mirror
    (an ABI drift here corrupts every call silently)."""
    import ctypes as C
    import subprocess
    from openimpala_b200 import capi
    fields_p = [f[0] for f in capi.oi_params._fields_]
    fields_i = [f[0] for f in capi.oi_solve_info._fields_]
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "openimpala_b200.h"', 'int main(void) {',
             '  printf("%zu %zu\\n", sizeof(oi_params), sizeof(oi_solve_info));']
    for f in fields_p:
        lines.append(f'  printf("p {f} %zu\\n", offsetof(oi_params, {f}));')
    for f in fields_i:
        lines.append(f'  printf("i {f} %zu\\n", offsetof(oi_solve_info, {f}));')
    lines += ['  return 0;', '}']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines()
    assert [int(v) for v in out[0].split()] == [C.sizeof(capi.oi_params), C.sizeof(capi.oi_solve_info)]
    for line in out[1:]:
        kind, name, off = line.split()
        st = capi.oi_params if kind == "p" else capi.oi_solve_info
        assert getattr(st, name).offset == int(off), (kind, name)
    # and the header compiles as plain C (extern "C" boundary, no C++ types)
    assert "oi_comm" in open(os.path.join(ROOT, "include", "openimpala_b200.h")).read()


def test_default_params_match_the_reference_defaults(built_lib):
    from openimpala_b200 import capi
    p = capi.default_params()
    assert (p.eps, p.maxiter) == (1e-9, 200)                 # TortuosityHypre.cpp:142-143
    assert (p.vlo, p.vhi) == (0.0, 1.0)                      # TortuosityHypre.H:77-78
    assert tuple(p.dx) == (1.0, 1.0, 1.0) and p.phase_id == 1 and p.direction == 0
    assert p.problem == capi.OI_PROBLEM_TORTUOSITY and p.halo_mode == capi.OI_HALO_AUTO and not p.comm


def test_argument_validation_needs_no_gpu(built_lib):
    """Bad arguments are rejected with OI_ERR_INVALID (-1) and a message, GPU or not."""
    import ctypes as C
    from openimpala_b200 import capi
    lib = built_lib
    h = C.c_void_p(None)
    for mutate, needle in [(lambda p: setattr(p, "nx", 0), "dimensions"), (lambda p: setattr(p, "direction", 3), "direction"),
                           (lambda p: setattr(p, "eps", 0.0), "eps"), (lambda p: setattr(p, "maxiter", 0), "iterations")]:
        p = capi.default_params()
        p.nx = p.ny = p.nz = 8
        mutate(p)
        assert lib.oi_create(C.byref(h), C.byref(p)) == -1 and not h.value
        assert needle in lib.oi_last_error().decode()
    assert lib.oi_create(None, None) == -1
    assert lib.oi_solve(None, None) == -1 and lib.oi_build_mask(None, None) == -1
    assert lib.oi_destroy(None) == 0                         # destroying nothing is fine
    pc, tc = C.c_int64(0), C.c_int64(0)
    assert lib.oi_count_phase_u8(None, 5, 1, C.byref(pc), C.byref(tc)) == -1
    assert lib.oi_count_phase_u8(None, -1, 1, C.byref(pc), C.byref(tc)) == -1


def test_sass_uses_blackwell_packed_fp32_and_async_copies(built_lib):
    """The level-0 kernels stage planes with cp.async (LDGSTS) and the two-sweeps-per-pass smoother
    does its stencil arithmetic with packed fp32 instructions (FFMA2 / FADD2 / FMUL2: sm_100 only)."""
    import shutil
    import subprocess
    from openimpala_b200 import capi
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert sass.count("LDGSTS") > 50
    assert sass.count("FFMA2") >= 10 and sass.count("FADD2") >= 6 and sass.count("FMUL2") >= 2
    assert "HMMA" not in sass and "UTCHMMA" not in sass          # no tensor-core detour: the path is HBM-bound


# ------------------------------------------------------------------ coarse tail on the host
def _tail_levels(rng, dims0, factors, periodic, alive_frac=0.8):
    """Random 7-point hierarchy: per level couplings to +x/+y/+z between live cells,
    diagonal = adjacent couplings + a sink, 0 on dead cells."""
    levels = []
    nx, ny, nz = dims0
    for (fx, fy, fz) in factors + [(1, 1, 1)]:
        alive = rng.random((nz, ny, nx)) < alive_frac
        c = []
        for axis in (2, 1, 0):
            nb = np.roll(alive, -1, axis)
            cc = np.where(alive & nb, 0.25 + rng.random((nz, ny, nx)), 0.0)
            if not periodic:
                idx = [slice(None)] * 3
                idx[axis] = -1
                cc[tuple(idx)] = 0.0
            c.append(cc.astype(np.float32))
        cxp, cyp, czp = c
        dg = np.zeros((nz, ny, nx))
        for axis, cc in ((2, cxp), (1, cyp), (0, czp)):
            dg += cc + np.roll(cc, 1, axis)
        dg = np.where(alive, dg + 0.05 + 0.3 * rng.random((nz, ny, nx)), 0.0).astype(np.float32)
        levels.append(dict(nx=nx, ny=ny, nz=nz, fx=fx, fy=fy, fz=fz, cxp=cxp, cyp=cyp, czp=czp, dg=dg))
        nx, ny, nz = -(-nx // fx), -(-ny // fy), -(-nz // fz)
    return levels


def _tail_reference_cycle(levels, l, b, w, wc):
    """The V-cycle of coarse_cycle() (oi_solver.cu) from its definition, in float64."""
    L = levels[l]
    dg = L["dg"].astype(np.float64)
    live = dg > 0

    def A(x):
        acc = dg * x
        for axis, key in ((2, "cxp"), (1, "cyp"), (0, "czp")):
            cp = L[key].astype(np.float64)
            acc = acc - cp * np.roll(x, -1, axis) - np.roll(cp, 1, axis) * np.roll(x, 1, axis)
        return acc

    def smooth(x, wt):
        return np.where(live, x + wt * (b - A(x)) / np.where(live, dg, 1.0), 0.0)

    last = (l + 1 == len(levels))
    ws = wc if last else w
    x = np.where(live, ws[0] * b / np.where(live, dg, 1.0), 0.0)
    for s in range(1, len(ws)):
        x = smooth(x, ws[s])
    if last:
        return x
    r = np.where(live, b - A(x), 0.0)
    C = levels[l + 1]
    fx, fy, fz = L["fx"], L["fy"], L["fz"]
    pad = np.zeros((C["nz"] * fz, C["ny"] * fy, C["nx"] * fx))
    pad[:L["nz"], :L["ny"], :L["nx"]] = r
    bc = pad.reshape(C["nz"], fz, C["ny"], fy, C["nx"], fx).sum(axis=(1, 3, 5))
    ec = _tail_reference_cycle(levels, l + 1, bc, w, wc)
    up = np.repeat(np.repeat(np.repeat(ec, fz, 0), fy, 1), fx, 2)[:L["nz"], :L["ny"], :L["nx"]]
    x = np.where(live, x + up, x)
    for s in range(len(w)):
        x = smooth(x, w[len(w) - 1 - s])
    return x


@pytest.fixture(scope="module")
def tail_emul(tmp_path_factory):
    """tests/cpu_emul/tail_emul.cu: the CUDA kernel's own tail_cycle() compiled for the host."""
    import ctypes
    import subprocess
    from openimpala_b200 import build
    out = tmp_path_factory.mktemp("tail_emul") / "libtail_emul.so"
    src = os.path.join(ROOT, "tests", "cpu_emul", "tail_emul.cu")
    subprocess.run([build.NVCC, "-O2", "-std=c++17", "--expt-relaxed-constexpr"] + build.ARCH +
                   ["-shared", "-Xcompiler", "-fPIC", "-cudart", "static", "-o", str(out), src], check=True)
    return ctypes.CDLL(str(out))


@pytest.mark.parametrize("dims0,factors,periodic,deg", [
    ((16, 16, 16), [(2, 2, 2), (2, 2, 2)], 0, 4),              # 16^3 -> 8^3 -> 4^3 (the 1024^3 tail)
    ((13, 13, 13), [(2, 2, 2), (2, 2, 2)], 0, 4),              # odd extents (the sample image's tail)
    ((12, 10, 14), [(2, 2, 2), (2, 2, 2)], 7, 4),              # periodic box (cell problem)
    ((9, 12, 20), [(1, 1, 2), (1, 2, 2), (2, 2, 2)], 0, 3),    # semicoarsened levels, odd degree
    ((6, 5, 4), [], 0, 4),                                     # the coarsest level alone
    ((7, 6, 5), [(2, 2, 1)], 7, 2),
])
@pytest.mark.parametrize("staged", [0, 1])
def test_coarse_tail_cycle_on_the_host(tail_emul, dims0, factors, periodic, deg, staged):
    """The one-CTA coarse tail (oi_coarse_tail.cuh) executed on the host through the same
    tail_cycle() template the kernel instantiates, against an independent numpy V-cycle:
    sweep order, weights, residual, aggregation restriction, piecewise-constant
    prolongation, periodic wrap, semicoarsening factors, result left in the first level's x."""
    import ctypes
    rng = np.random.default_rng(hash((dims0, periodic, deg)) % (2 ** 32))
    levels = _tail_levels(rng, dims0, list(factors), bool(periodic))
    w = [1.0 / (0.3 + 0.4 * k) for k in range(deg)]
    wc = [1.0 / (0.2 + 0.22 * k) for k in range(8)]
    b0 = np.where(levels[0]["dg"] > 0, rng.standard_normal(levels[0]["dg"].shape), 0.0).astype(np.float32)
    ref = _tail_reference_cycle(levels, 0, b0.astype(np.float64), w, wc)
    dims, fields, keep = [], [], []
    for q, L in enumerate(levels):
        dims += [L["nx"], L["ny"], L["nz"], L["fx"], L["fy"], L["fz"]]
        n = L["nx"] * L["ny"] * L["nz"]
        arrs = [np.ascontiguousarray(L[k]).reshape(n) for k in ("cxp", "cyp", "czp", "dg")]
        x = rng.standard_normal(n).astype(np.float32)          # garbage: the cycle starts from zero itself
        b = b0.reshape(n).copy() if q == 0 else rng.standard_normal(n).astype(np.float32)
        t = rng.standard_normal(n).astype(np.float32)
        arrs += [x, b, t]
        keep.append(arrs)
        fields += [a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) for a in arrs]
    c_dims = (ctypes.c_int * len(dims))(*dims)
    c_fields = (ctypes.POINTER(ctypes.c_float) * len(fields))(*fields)
    c_w = (ctypes.c_double * deg)(*w)
    c_wc = (ctypes.c_double * 8)(*wc)
    before = [[a.copy() for a in arrs] for arrs in keep]
    rc = tail_emul.oi_tail_emulate(len(levels), c_dims, int(periodic), c_fields, deg, c_w, 8, c_wc, int(staged))
    assert rc == 0
    if staged:      # the staged variant touches global memory only to read inputs and to write the first level's x
        for q, (arrs, old) in enumerate(zip(keep, before)):
            for m, (a, o) in enumerate(zip(arrs, old)):
                if not (q == 0 and m == 4):
                    assert np.array_equal(a, o), (q, m)
    got = keep[0][4].reshape(ref.shape)
    scale = float(np.abs(ref).max())
    assert scale > 0
    assert float(np.abs(got - ref).max()) <= 2e-5 * scale       # fp32 cycle against the float64 reference
    assert not got[levels[0]["dg"] == 0].any()                  # empty aggregates stay zero


# ------------------------------------------------------------------ AMReX stand-in bulk routines
def test_amrex_shim_bulk_routines(tmp_path):
    """FillBoundary (ghost shell only), Copy (row copies) and min / max / sum of the AMReX
    stand-in against naive per-cell loops: 72 cases over ghost widths, periodic flags, a
    non-zero lower corner and degenerate extents (tests/cpu_emul/shim_check.cpp)."""
    import subprocess
    exe = tmp_path / "shim_check"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "openimpala_b200", "host", "amrex_shim"),
                    "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpu_emul", "shim_check.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and "mismatches 0" in r.stdout, r.stdout + r.stderr


def test_cpp_slab_rule_equals_python_slab_partition(tmp_path):
    """The C++ stand-in's ParallelDescriptor::slab (which z-slab a rank of `Diffusion b200.ranks=N` owns) and
    capi.slab_partition (bench.py, the Python classes) must cut the box the same way: multigrid aggregates never
    straddle ranks only if both sides align the slab boundaries identically."""
    import subprocess
    from openimpala_b200 import capi
    src = tmp_path / "slab.cpp"
    src.write_text('#include <AMReX.H>\n#include <cstdio>\n#include <cstdlib>\n'
                   'int main(int c, char** v) { int nz = atoi(v[1]), n = atoi(v[2]);\n'
                   '  for (int r = 0; r < n; ++r) { int z0, nzl; amrex::ParallelDescriptor::slab(nz, n, r, z0, nzl); printf("%d %d\\n", z0, nzl); } }\n')
    exe = tmp_path / "slab"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "openimpala_b200", "host", "amrex_shim"),
                    "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    for nz, n in [(100, 2), (100, 4), (1024, 8), (1024, 2), (1280, 2), (1536, 4), (2048, 8), (64, 8), (97, 3), (512, 4), (130, 5)]:
        out = subprocess.run([str(exe), str(nz), str(n)], capture_output=True, text=True, check=True).stdout.split()
        got = [(int(out[2 * r]), int(out[2 * r + 1])) for r in range(n)]
        assert got == capi.slab_partition(nz, n), (nz, n, got)


def test_cpp_fields_hold_the_rank_slab_and_leave_z_ghosts_to_the_neighbours(tmp_path):
    """SPMD stand-in (`OI_RANK` / `OI_WORLD_SIZE`): a field of a 2-rank run holds this rank's z-slab of the BoxArray,
    `validCopy` is that slab, and FillBoundary wraps the periodic x / y ghosts inside the slab but leaves the z ghost
    planes alone -- they belong to the neighbour ranks and are the library's to exchange (the homogenisation method
    of `Diffusion b200.ranks=N` relies on exactly this)."""
    import subprocess
    from openimpala_b200 import capi
    src = tmp_path / "slabfab.cpp"
    src.write_text(r"""
#include <AMReX.H>
#include <cstdio>
int main() {
    using namespace amrex;
    Box dom(IntVect(0, 0, 0), IntVect(7, 5, 99));
    BoxArray ba(dom);
    DistributionMapping dm(ba);
    iMultiFab f(ba, dm, 1, 1);
    const Box loc = ba.localBox();
    f.setVal(-7);
    for (int k = loc.smallEnd(2); k <= loc.bigEnd(2); ++k)
        for (int j = 0; j <= 5; ++j)
            for (int i = 0; i <= 7; ++i) f(i, j, k, 0) = i + 10 * j + 100 * k;
    Periodicity per; per.p = {1, 1, 1};
    f.FillBoundary(per);
    int bad = 0;
    const int k0 = loc.smallEnd(2), k1 = loc.bigEnd(2);
    for (int k = k0; k <= k1; ++k)
        for (int j = -1; j <= 6; ++j)
            for (int i = -1; i <= 8; ++i) {
                const int want = (i + 8) % 8 + 10 * ((j + 6) % 6) + 100 * k;
                if (f(i, j, k, 0) != want) ++bad;
            }
    for (int j = -1; j <= 6; ++j)
        for (int i = -1; i <= 8; ++i) {
            if (f(i, j, k0 - 1, 0) != -7) ++bad;            // z ghosts untouched
            if (f(i, j, k1 + 1, 0) != -7) ++bad;
        }
    const std::vector<int> v = f.validCopy(0);
    if ((long long)v.size() != loc.numPts() || v.front() != 100 * k0 || v.back() != 7 + 50 + 100 * k1) ++bad;
    std::printf("%d %d %d\n", k0, loc.length(2), bad);
    return 0;
}
""")
    exe = tmp_path / "slabfab"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I" + os.path.join(ROOT, "openimpala_b200", "host", "amrex_shim"),
                    "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    parts = capi.slab_partition(100, 2)
    for rank in (0, 1):
        out = subprocess.run([str(exe)], capture_output=True, text=True, check=True,
                             env={**os.environ, "OI_RANK": str(rank), "OI_WORLD_SIZE": "2"}).stdout.split()
        assert (int(out[0]), int(out[1])) == parts[rank] and int(out[2]) == 0, (rank, out)
    # one rank: the whole box, z ghosts wrapped as well
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True,
                         env={k: v for k, v in os.environ.items() if k not in ("OI_RANK", "OI_WORLD_SIZE", "RANK", "WORLD_SIZE")}
                         ).stdout.split()
    assert (int(out[0]), int(out[1])) == (0, 100) and int(out[2]) == 2 * 8 * 10, out     # the two wrapped z ghost planes
