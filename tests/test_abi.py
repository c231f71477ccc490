"""The C-ABI shared library loads (no GPU needed) and exports every symbol that
include/rawspeed_b200.h declares; struct layouts match the ctypes mirrors."""
import ctypes as C
import os
import re
import subprocess

from rawspeed_b200 import _abi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "rawspeed_b200.h")


def declared_functions():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rsb200_[a-z0-9_]+)\s*\(", src)))


def test_header_functions_are_exported():
    lib = _abi.load()
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(_abi.EXPORTS)
    assert lib.rsb200_abi_version() == 1


def test_struct_layouts_match_header(tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof
    with the ctypes mirrors used by the Python side."""
    prog = tmp_path / "layout.c"
    prog.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "rawspeed_b200.h"
int main(void){
  printf("%zu %zu %zu %zu %zu\n", sizeof(rsb200_unpack_job), sizeof(rsb200_huff_table),
         sizeof(rsb200_ljpeg_scan), sizeof(rsb200_scan_result), sizeof(rsb200_cr2_job));
  printf("%zu %zu %zu\n", sizeof(rsb200_pentax_job), offsetof(rsb200_pentax_job, width), offsetof(rsb200_pentax_job, out_pitch));
  printf("%zu %zu %zu %zu\n", sizeof(rsb200_sraw_job), offsetof(rsb200_sraw_job, sraw_coeffs), offsetof(rsb200_sraw_job, out_offset), offsetof(rsb200_sraw_job, out_pitch));
  printf("%zu %zu %zu\n", sizeof(rsb200_raw_job), offsetof(rsb200_raw_job, format), offsetof(rsb200_raw_job, table));
  printf("%zu %zu %zu %zu\n", offsetof(rsb200_ljpeg_scan, init_pred), offsetof(rsb200_ljpeg_scan, out_offset),
         offsetof(rsb200_cr2_job, frame_w), offsetof(rsb200_cr2_job, out_offset));
  printf("%zu %zu %zu %zu\n", sizeof(rsb200_pana_job), offsetof(rsb200_pana_job, version),
         offsetof(rsb200_pana_job, zero_is_not_bad), offsetof(rsb200_pana_job, section_split_offset));
  printf("%zu %zu %zu %zu\n", sizeof(rsb200_scale_job), offsetof(rsb200_scale_job, black_separate),
         offsetof(rsb200_scale_job, white_point), offsetof(rsb200_scale_job, path));
  printf("%zu %zu %zu %zu %zu\n", sizeof(rsb200_arw2_job), sizeof(rsb200_nikon_job),
         sizeof(rsb200_phaseone_job), sizeof(rsb200_phaseone_strip), offsetof(rsb200_nikon_job, pup));
  printf("%zu %zu\n", sizeof(rsb200_lookup_job), offsetof(rsb200_lookup_job, table));
  printf("%zu %zu %zu %zu %zu\n", sizeof(rsb200_dng_op), sizeof(rsb200_dngop_job), sizeof(rsb200_badpix_job),
         offsetof(rsb200_badpix_job, prior_map), offsetof(rsb200_dng_op, value));
  return 0; }''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    got = [int(x) for x in out]
    want = [C.sizeof(_abi.UnpackJob), C.sizeof(_abi.HuffTable), C.sizeof(_abi.LJpegScan),
            C.sizeof(_abi.ScanResult), C.sizeof(_abi.Cr2Job),
            C.sizeof(_abi.PentaxJob), _abi.PentaxJob.width.offset, _abi.PentaxJob.out_pitch.offset,
            C.sizeof(_abi.SrawJob), _abi.SrawJob.sraw_coeffs.offset,
            _abi.SrawJob.out_offset.offset, _abi.SrawJob.out_pitch.offset,
            C.sizeof(_abi.RawJob), _abi.RawJob.format.offset, _abi.RawJob.table.offset,
            _abi.LJpegScan.init_pred.offset, _abi.LJpegScan.out_offset.offset,
            _abi.Cr2Job.frame_w.offset, _abi.Cr2Job.out_offset.offset,
            C.sizeof(_abi.PanaJob), _abi.PanaJob.version.offset, _abi.PanaJob.zero_is_not_bad.offset,
            _abi.PanaJob.section_split_offset.offset,
            C.sizeof(_abi.ScaleJob), _abi.ScaleJob.black_separate.offset,
            _abi.ScaleJob.white_point.offset, _abi.ScaleJob.path.offset,
            C.sizeof(_abi.Arw2Job), C.sizeof(_abi.NikonJob), C.sizeof(_abi.PhaseOneJob),
            C.sizeof(_abi.PhaseOneStrip), _abi.NikonJob.pup.offset,
            C.sizeof(_abi.LookupJob), _abi.LookupJob.table.offset,
            C.sizeof(_abi.DngOp), C.sizeof(_abi.DngOpJob), C.sizeof(_abi.BadPixJob),
            _abi.BadPixJob.prior_map.offset, _abi.DngOp.value.offset]
    assert got == want


def test_no_gpu_means_loud_failure_not_fallback():
    """Without a usable device the product raises (this container has no GPU);
    with one it must construct.  Either way there is no CPU path."""
    import rawspeed_b200 as rs
    import torch
    if torch.cuda.is_available():
        rs.Context(0).close()
    else:
        try:
            rs.Context(0)
        except rs.Rsb200Error as e:
            assert e.code == _abi.ERR_CUDA
        else:
            raise AssertionError("Context() succeeded without a GPU")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rawspeed_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "rs_oracle" not in txt and "import oracle" not in txt and \
                    "from oracle" not in txt and "libref" not in txt, (dirpath, f)
    assert os.path.exists(build.LIB)
