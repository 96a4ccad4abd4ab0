"""CPU tests of the drop-in boundary: libzvdb_b200.so loads, exports every symbol that
include/zvdb_b200.h declares, and refuses to work (loudly) without a GPU. No compute calls."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "zvdb_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"ZVDB_API[^;(]*?\b(zvdb_\w+)\s*\(", txt)))


def test_header_declares_the_boundary():
    names = _declared()
    for must in ("zvdb_create", "zvdb_destroy", "zvdb_insert", "zvdb_search", "zvdb_search_batch",
                 "zvdb_search_batch_device", "zvdb_count", "zvdb_get_point", "zvdb_load_graph",
                 "zvdb_merge_topk_device", "zvdb_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol(zv):
    lib = C.CDLL(zv._lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/zvdb_b200.h but not exported"


def test_python_binding_covers_the_header(zv):
    assert sorted(zv._lib.SIGNATURES) == _declared()


def test_version_and_error_strings(zv):
    assert b"sm_100a" in zv.lib().zvdb_version()
    assert zv.lib().zvdb_last_error() is not None


def test_library_holds_sm100a_code_only(zv):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", zv._lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\w+", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback(zv):
    """Without a CUDA device every entry point that would compute fails with ZVDB_ERR_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is only reachable without one")
    with pytest.raises(zv.ZvdbError) as e:
        zv.HNSW(16, 200)
    assert e.value.code == zv._lib.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under zvdb_b200/ may reference it."""
    pkg = os.path.join(ROOT, "zvdb_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
                assert "orc_" not in txt, f


def test_header_is_plain_c_and_a_c_caller_links(zv, tmp_path):
    """The boundary is a C ABI: include/zvdb_b200.h must compile as C (no C++ / torch types) and a C program must
    link against libzvdb_b200.so and reach an entry point. Without a GPU the create call fails loudly (no CPU path)."""
    import os
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include <stdio.h>
#include "zvdb_b200.h"
int main(void) {
    zvdb_index *ix = 0;
    int rc = zvdb_create(&ix, 0, 16, 200, 0, 0);
    if (rc != 0) { printf("create failed rc=%d: %s\n", rc, zvdb_last_error()); return 3; }
    float p[4] = {1.f, 2.f, 3.f, 4.f}, q[4] = {1.f, 2.f, 3.f, 5.f}, d[1]; uint64_t id[1]; uint32_t cnt = 0;
    if (zvdb_insert(ix, p, 4) != 0 || zvdb_search(ix, q, 4, 1, id, d, &cnt) != 0) { printf("%s\n", zvdb_last_error()); return 4; }
    printf("ok id=%llu d=%g n=%u\n", (unsigned long long)id[0], d[0], cnt);
    zvdb_destroy(ix);
    return 0;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.join(root, "zvdb_b200", "lib")
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                        "-L", libdir, "-lzvdb_b200", f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([str(exe)], capture_output=True, text=True)
    import torch
    if torch.cuda.is_available():
        assert run.returncode == 0 and run.stdout.startswith("ok id=0 d=1 n=1"), run.stdout + run.stderr
    else:
        assert run.returncode == 3 and "create failed" in run.stdout, run.stdout + run.stderr


def test_search_kernels_fit_their_register_budget_without_spills(zv):
    """Every instantiation of the hot-path kernel must fit the register budget its residency needs (one-warp CTAs:
    32 per SM -> 64 registers for rows up to 1 KiB) without spills in the pop loop: a spill there cost 12-19 % when
    it was measured (profiles/r01_k1_experiments.md), so it must not come back unnoticed. Of the plain-search
    instantiations (EXCH = false) every shared-memory-hash one (the mode of every default bench line) and the 128-d L2
    bitmap one (C2's large-ef mode) have no stack frame at all; the sharded-step instantiations and the other global-visited ones (256-d rows, cosine / dot, the hash fallback for > 14.5 M-row shards)
    keep one 4-byte value (the lane id) on the stack, re-read only in the per-query epilogue (checked in the SASS when it
    appeared, round 2) -- 8 bytes of stack are tolerated there and nowhere else."""
    import re
    import shutil
    import subprocess
    from zvdb_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([cuobjdump, "-res-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    rows = re.findall(r"Function (\S*search_layer0_kernel\S*):\s*\n\s*REG:(\d+) STACK:(\d+)", out)
    assert len(rows) >= 90, "search_layer0_kernel instantiations not found in the library"
    for name, reg, stack in rows:
        m = re.search(r"search_layer0_kernelILi(\d+)ELi(\d)ELi(\d)ELb([01])E", name)       # <CPL, METRIC, VIS, EXCH>
        assert m, name
        cpl, metric, vis, exch = int(m.group(1)), int(m.group(2)), int(m.group(3)), int(m.group(4))
        allowed = 0 if (not exch and (vis == 0 or (cpl == 1 and metric == 0 and vis == 1))) else 8
        assert int(stack) <= allowed, f"{name} spills ({stack} bytes of stack)"
        if cpl <= 2:
            assert int(reg) <= 64, f"{name}: {reg} registers, 32 one-warp CTAs per SM need <= 64"


def test_team_kernel_fits_its_residency_without_spills(zv):
    """K1L (one CTA of 256 threads per query, search_team_kernel.cuh): no instantiation spills, and the 128-d ones fit the
    register budget the launch rule counts on (capi.cu launch_search: three 256-thread or six 128-thread CTAs per SM at one
    16-byte chunk pair per lane -> 768 threads x <= 80 registers; two / four per SM up to 512-d -> <= 128). The m = 16 instantiations with the
    on-chip adjacency cache copy adjacency rows with cp.async (LDGSTS in the SASS)."""
    import re
    import shutil
    import subprocess
    from zvdb_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    out = subprocess.run([cuobjdump, "-res-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    rows = re.findall(r"Function (\S*search_team_kernel\S*):\s*\n\s*REG:(\d+) STACK:(\d+)", out)
    assert len(rows) == 5 * 3 * 4, f"{len(rows)} search_team_kernel instantiations (expected CPL x METRIC x {{plain, cached, cached m=16, half team}})"
    for name, reg, stack in rows:
        m = re.search(r"search_team_kernelILi(\d+)ELi(\d)ELb([01])ELi(\d+)ELi(\d+)E", name)      # <CPL, METRIC, ADJC, MC, T>
        assert m, name
        cpl, threads = int(m.group(1)), int(m.group(5))
        assert threads in (128, 256), name
        assert int(stack) == 0, f"{name} spills ({stack} bytes of stack)"
        assert int(reg) <= (80 if cpl == 1 else 128 if cpl <= 4 else 255), f"{name}: {reg} registers"
    sass = subprocess.run([cuobjdump, "-sass", "-fun", "_ZN4zvdb18search_team_kernelILi1ELi0ELb1ELi16ELi256EEEvNS_12SearchParamsE", _lib.LIB_PATH],
                          capture_output=True, text=True).stdout
    assert "LDGSTS" in sass and "CREDUX" in sass and "FFMA2" in sass, "K1L m=16: cp.async / REDUX / packed FMA not found"


def test_library_sass_holds_the_blackwell_instructions_the_design_claims(zv):
    """DESIGN.md's claims about the machine code, checked on the built library: the exact k-NN kernel issues 5th-gen
    tensor-core MMAs on CTA pairs with TMA loads and TMEM reads (UTCHMMA.2CTA, UTMALDG.2D.2CTA, UTCBAR multicast commits,
    LDTM), the search kernel uses packed f32x2 FMAs, warp REDUX for the pop selection and L2 prefetches."""
    import shutil
    import subprocess
    from zvdb_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    sass = subprocess.run([cuobjdump, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA.2CTA", "UTMALDG.2D.2CTA", "UTCBAR.2CTA.MULTICAST", "LDTM.x32", "SYNCS.PHASECHK.TRANS64.TRYWAIT",
                     "FFMA2", "REDUX", "CCTL.E.PF2"):
        assert mnemonic in sass, f"{mnemonic} not found in libzvdb_b200.so"
    assert "HMMA.16" not in sass and "WGMMA" not in sass     # no mma.sync / wgmma fallback kernels


def test_c_benchmark_harness_compiles_and_refuses_without_a_gpu(zv, tmp_path):
    """integration/harness.c (the reference's benchmark loops over the C ABI, for per-call numbers without the Python
    interpreter) builds warning-free as C99 against the header and the library; without a GPU it fails loudly."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc:
        pytest.skip("no C compiler")
    libdir = os.path.join(ROOT, "zvdb_b200", "lib")
    exe = tmp_path / "harness"
    r = subprocess.run([cc, "-O2", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "integration", "harness.c"), "-o", str(exe), "-L", libdir, "-lzvdb_b200",
                        f"-Wl,-rpath,{libdir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([str(exe), "200", "16", "50", "5"], capture_output=True, text=True)
    import torch
    if torch.cuda.is_available():
        assert run.returncode == 0 and "Search per second" in run.stdout, run.stdout + run.stderr
    else:
        assert run.returncode == 1 and "no CPU fallback" in run.stderr, run.stdout + run.stderr
