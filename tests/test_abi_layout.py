"""The ctypes mirrors of include/b200fft.h (mpifft4py_b200/_cdefs.py) against the C compiler's view of the
structs: total size and the offset of every field.  A field added on one side only would shift everything
behind it silently."""
import ctypes as C
import os
import subprocess
import tempfile

from mpifft4py_b200 import _cdefs as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAIRS = [("b200fft_ns_mesh_t", D.NsMesh, {}), ("b200fft_side_t", D.Side, {}), ("b200fft_mask_t", D.Mask, {}),
         ("b200fft_strided_desc_t", D.StridedDesc, {"inp": "in"}), ("b200fft_rows_desc_t", D.RowsDesc, {}),
         ("b200fft_plan_desc_t", D.PlanDesc, {})]


def test_struct_layouts_match_the_header():
    lines = ['#include <cstdio>', '#include <cstddef>', '#include "include/b200fft.h"', 'int main() {']
    expected = []
    for cname, ct, rename in PAIRS:
        lines.append('  std::printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        expected.append("%s %d" % (cname, C.sizeof(ct)))
        for fname, _ in ct._fields_:
            cf = rename.get(fname, fname)
            lines.append('  std::printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, cf, cname, cf))
            expected.append("%s.%s %d" % (cname, cf, getattr(ct, fname).offset))
    lines += ['  return 0;', '}']
    with tempfile.TemporaryDirectory() as tmp:
        src, exe = os.path.join(tmp, "layout.cpp"), os.path.join(tmp, "layout")
        open(src, "w").write("\n".join(lines))
        subprocess.run(["g++", "-I", ROOT, src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, stdout=subprocess.PIPE).stdout.decode().split("\n")
    assert [ln for ln in out if ln] == expected


def test_integration_md_stub_has_the_same_structs():
    """The ctypes stub INTEGRATION.md shows a maintainer (section B.1) declares the structs field for field like
    _cdefs.py: a stale stub would hand the library a mis-sized descriptor."""
    import re
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    ns = {"C": C, "MAXP": D.MAXP}
    found = {}
    for m in re.finditer(r"^class (\w+)\(C\.Structure\):[^\n]*\n((?:[ \t]+[^\n]*\n)+)", text, re.M):
        exec("class %s(C.Structure):\n%s" % (m.group(1), m.group(2)), ns)
        found[m.group(1)] = ns[m.group(1)]
    assert set(found) >= {"Side", "Mask", "StridedDesc"}
    for name, ct in found.items():
        mine = getattr(D, name)
        assert [f[0] for f in ct._fields_] == [f[0] for f in mine._fields_], name
        assert C.sizeof(ct) == C.sizeof(mine), name
        for f in ct._fields_:
            assert getattr(ct, f[0]).offset == getattr(mine, f[0]).offset, (name, f[0])
