"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/desamba_b200.h declares,
record layouts match the header, and without a GPU the compute entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "desamba_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dsb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import desamba_b200 as dsb
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(dsb.lib, n), f"{n} declared in include/desamba_b200.h but not exported"


def test_record_layouts_match_header():
    # compile a tiny C program against the header and compare sizeof/offsetof with the numpy dtypes of the binding
    import desamba_b200 as dsb
    prog = r'''
#include "desamba_b200.h"
#include <stdio.h>
#include <stddef.h>
int main(void){
 printf("%zu %zu %zu %zu\n", sizeof(dsb_hit), sizeof(dsb_read_result), sizeof(dsb_seed), sizeof(dsb_ref_info));
 printf("%zu %zu %zu %zu\n", offsetof(dsb_hit, direction), offsetof(dsb_read_result, n_anchor), offsetof(dsb_read_result, read_len), offsetof(dsb_ref_info, seq_offset));
 return 0; }'''
    exe = "/tmp/dsb_layout_test"
    subprocess.run(["gcc", "-x", "c", "-", "-I", os.path.join(ROOT, "include"), "-o", exe], input=prog.encode(), check=True)
    out = subprocess.run([exe], capture_output=True, check=True).stdout.split()
    assert [int(x) for x in out[:4]] == [dsb.HIT_DTYPE.itemsize, dsb.RR_DTYPE.itemsize, dsb.SEED_DTYPE.itemsize, 144]
    assert int(out[4]) == dsb.HIT_DTYPE.fields["direction"][1]
    assert int(out[5]) == dsb.RR_DTYPE.fields["n_anchor"][1]
    assert int(out[6]) == dsb.RR_DTYPE.fields["read_len"][1]
    assert int(out[7]) == 136


def test_no_gpu_fails_loudly(tmp_path):
    import torch
    import desamba_b200 as dsb
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path cannot be exercised")
    with pytest.raises(dsb.DsbError) as e:
        dsb.Index(str(tmp_path), 0)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)


def test_bad_arguments_return_codes():
    import desamba_b200 as dsb
    assert dsb.lib.dsb_index_load(None, 0, None) == -1          # DSB_E_ARG, never an abort across the ABI
    assert dsb.lib.dsb_batch_run(None, 0) == -1
    assert dsb.lib.dsb_batch_launches(None) == 0
    o = dsb.api.Opts()
    dsb.lib.dsb_opts_default(C.byref(o))
    assert (o.l_min_match, o.min_score) == (170, 64)            # cly_mt.c:486


def test_driver_binary_usage():
    exe = os.path.join(ROOT, "desamba_b200", "bin", "deSAMBA-b200")
    r = subprocess.run([exe, "classify"], capture_output=True)
    assert r.returncode == 0 and b"Usage" in r.stderr            # like the reference: usage + exit 0 (cly_mt.c:504-508)
    r = subprocess.run([exe, "index"], capture_output=True)
    assert r.returncode == 1


def test_product_does_not_touch_the_oracle():
    # the product path must not import, link or execute anything under oracle/
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "desamba_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h", "Makefile")):
                txt = open(os.path.join(d, f), errors="replace").read()
                if re.search(r"oracle/|liborc|desamba_oracle|orc_capi", txt):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
    out = subprocess.run(["ldd", os.path.join(ROOT, "desamba_b200", "lib", "libdesamba_b200.so")], capture_output=True).stdout
    assert b"liborc" not in out
