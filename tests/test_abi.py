"""CPU tests: the C-ABI library loads and exports every symbol include/spmv_b200.h declares (no compute calls)."""
import ctypes
import re
from pathlib import Path

from spmv_acc_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def _declared_symbols():
    text = (ROOT / "include" / "spmv_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spmv_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_list_agree():
    assert _declared_symbols() == sorted(_lib.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    so = _lib.LIB_DIR / "libspmv_b200.so"
    assert so.exists(), "run `python -m spmv_acc_b200.build` first"
    L = ctypes.CDLL(str(so))
    for s in _declared_symbols():
        assert hasattr(L, s), f"{s} is declared in include/spmv_b200.h but not exported"
    L.spmv_b200_abi_version.restype = ctypes.c_int
    assert L.spmv_b200_abi_version() == 2


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_lib.Options) == 20
    # int32 m,n | int64 nnz | 4 x int32 | uint32 | 2 x int32 | 3 x int32 | 3 x int32 (+4 pad) | 8 x int64 | 2 x int64
    # | 2 x int64 | 4 x int32
    assert ctypes.sizeof(_lib.PlanInfo) == 8 + 8 + 16 + 4 + 8 + 12 + 16 + 64 + 16 + 16 + 16
    assert ctypes.sizeof(_lib.HaloLoopDesc) == 712 and ctypes.sizeof(_lib.HaloLoopInfo) == 24


def test_struct_layouts_match_the_c_compiler(tmp_path):
    """sizeof of every ABI struct as gcc sees include/spmv_b200.h equals the ctypes mirror."""
    import shutil
    import subprocess
    import pytest
    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    if not Path(gcc).exists():
        pytest.skip("gcc not available")
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "spmv_b200.h"\nint main(void) { printf("%zu %zu %zu %zu %zu\\n", '
                   'sizeof(spmv_b200_options), sizeof(spmv_b200_plan_info), sizeof(spmv_b200_push), '
                   'sizeof(spmv_b200_halo_loop_desc), sizeof(spmv_b200_halo_loop_info)); return 0; }\n')
    exe = tmp_path / "sz"
    r = subprocess.run([gcc, "-std=c99", f"-I{ROOT / 'include'}", str(src), "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    want = [ctypes.sizeof(t) for t in (_lib.Options, _lib.PlanInfo, _lib.Push, _lib.HaloLoopDesc, _lib.HaloLoopInfo)]
    assert got == want


def test_product_package_never_imports_the_oracle():
    for py in (ROOT / "spmv_acc_b200").rglob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f"{py} must not depend on oracle/"
    # the CUDA / C++ sources may cite the oracle in comments (which restatement mirrors an array), never use it: with
    # comments stripped the word must not occur, so nothing under oracle/ can be included, linked or dlopen'ed
    for src in list((ROOT / "spmv_acc_b200" / "csrc").glob("*")) + list((ROOT / "src" / "acc" / "cuda-b200").glob("*")) + \
            list((ROOT / "include").glob("*.h")):
        code = re.sub(r"/\*.*?\*/", "", src.read_text(), flags=re.S)
        code = re.sub(r"//[^\n]*", "", code)
        assert "oracle" not in code, f"{src} refers to the oracle outside a comment"
    build_py = (ROOT / "spmv_acc_b200" / "build.py").read_text()
    assert "oracle" not in build_py, "the product build must not compile or link anything under oracle/"


def test_sass_contains_tma_bulk_copies():
    """The streaming kernels are TMA kernels: UBLKCP must be present in the sm_100a SASS."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        import pytest
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-sass", str(_lib.LIB_DIR / "libspmv_b200.so")], capture_output=True, text=True)
    assert "UBLKCP" in out.stdout and "sm_100a" in out.stdout


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The boundary is a C ABI: include/spmv_b200.h must compile as C99 (no C++-isms) and a C program must link against
    the library and call an entry point that needs no GPU."""
    import shutil
    import subprocess
    import pytest
    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    if not Path(gcc).exists():
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include "spmv_b200.h"\n'
                   'int main(void) {\n'
                   '  spmv_b200_options o = {0, 0, 0, 0, SPMV_B200_FLAG_BETA0_SKIP_Y};\n'
                   '  spmv_b200_halo_loop_desc d; spmv_b200_push p; (void)o; (void)d; (void)p;\n'
                   '  printf("%d %d\\n", spmv_b200_abi_version(), spmv_b200_cache_size());\n'
                   '  return spmv_b200_plan_destroy(0);\n}\n')
    exe = tmp_path / "abi"
    lib_dir = _lib.LIB_DIR
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{ROOT / 'include'}", str(src),
                        "-o", str(exe), f"-L{lib_dir}", "-lspmv_b200", f"-Wl,-rpath,{lib_dir}"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split() == ["2", "0"], (out.stdout, out.stderr)


def test_direct_kernel_uses_no_shared_memory_and_lane_contiguous_stream_loads():
    """The direct form exists to leave the whole unified L1/shared array to L1: its kernels must not use shared memory or
    a CTA barrier. Element ownership is lane-contiguous (element 32j + lane per instruction), so a 128-element window is
    streamed with four 32-bit colindex and four 64-bit value loads per lane (evict-first, SASS LDG.E.EF[.64]), all of
    them issued before the first product, and x is gathered with 64-bit read-only loads; no fp64 atomics."""
    import shutil
    import subprocess
    import pytest
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    so = str(_lib.LIB_DIR / "libspmv_b200.so")
    import re
    res = subprocess.run([cuobjdump, "-res-usage", so], capture_output=True, text=True).stdout.splitlines()
    seen = 0
    for i, line in enumerate(res):
        if "k_spmv_warp" in line:
            seen += 1
            shared = int(re.search(r"SHARED:(\d+)", res[i + 1]).group(1))
            assert shared <= 1024, res[i + 1]  # 1024 bytes per CTA are reserved by the system on sm_100 for every kernel
            regs = int(re.search(r"REG:(\d+)", res[i + 1]).group(1))
            assert regs <= 40, res[i + 1]
    assert seen == 2  # register budgets for 8 and 6 resident CTAs per SM
    elf = subprocess.run([cuobjdump, "-elf", so], capture_output=True, text=True).stdout
    fn = re.search(r"_ZN4b200\d+k_spmv_warpILi8E\w*", elf).group(0)
    sass = subprocess.run([cuobjdump, "-sass", "-fun", fn, so], capture_output=True, text=True).stdout
    assert "BAR.SYNC" not in sass and "ATOM" not in sass and "RED." not in sass
    lines = [ln for ln in sass.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", ln)]
    first_dmul = next(i for i, ln in enumerate(lines) if "DMUL" in ln)
    head = "\n".join(lines[:first_dmul])
    assert len(re.findall(r"LDG\.E\.EF\.64", head)) >= 4      # value: element 32j + lane, j = 0..3
    assert len(re.findall(r"LDG\.E\.EF ", head)) >= 4          # colindex
    assert len(re.findall(r"LDG\.E\.64\.CONSTANT", head)) >= 4  # x gathers


def test_push_descriptor_marks_multicast_destinations():
    """(row_lo, row_hi, address[, is_multicast]) tuples -> spmv_b200_push: the mask bit of a destination is set only for
    entries flagged as NVLink multicast addresses; the plain three-field form keeps working."""
    from spmv_acc_b200 import api
    ps = api._fill_push(_lib.Push(), [(0, 10, 0x1000), (10, 20, 0x2000, True), (5, 6, 0x3000, False)])
    assert ps.count == 3 and ps.multicast_mask == 0b010
    assert [ps.row_lo[j] for j in range(3)] == [0, 10, 5] and [ps.row_hi[j] for j in range(3)] == [10, 20, 6]
    assert [ps.dst[j] for j in range(3)] == [0x1000, 0x2000, 0x3000]
    assert api._fill_push(_lib.Push(), []).count == 0
