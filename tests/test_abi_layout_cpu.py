"""The ctypes mirrors in nerf_lidar_b200/_lib.py against include/nlb200.h compiled by gcc (no GPU): size of every
struct and offset of every field.  The boundary is a C ABI; a host-side struct that drifts from the header passes
garbage to a kernel without any error."""
import ctypes as C
import os
import re
import shutil
import subprocess

import pytest

from nerf_lidar_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'nlb200.h')

MIRRORS = {
    'nlb_rays_t': 'NlbRays', 'nlb_table_t': 'NlbTable', 'nlb_losses_in_t': 'NlbLossesIn',
    'nlb_composite_in_t': 'NlbCompositeIn', 'nlb_composite_out_t': 'NlbCompositeOut',
    'nlb_composite_grad_t': 'NlbCompositeGrad', 'nlb_nerf_mlp_weights_t': 'NlbNerfMlpWeights',
    'nlb_nerf_mlp_wgrads_t': 'NlbNerfMlpWeights', 'nlb_nerf_mlp_saved_t': 'NlbNerfMlpSaved',
    'nlb_nerf_mlp_grad_in_t': 'NlbNerfMlpGradIn', 'nlb_nerf_mlp_grad_out_t': 'NlbNerfMlpGradOut',
    'nlb_bf16_sum_job_t': 'NlbBf16SumJob', 'nlb_sum_term_t': 'NlbSumTerm', 'nlb_scale_job_t': 'NlbScaleJob',
    'nlb_ray_grads_t': 'NlbRayGrads', 'nlb_ray_out_t': 'NlbRayOut', 'nlb_obj_mlp_t': 'NlbObjMlp',
    'nlb_obj_grads_t': 'NlbObjGrads', 'nlb_range_image_t': 'NlbRangeImage', 'nlb_unet_conv_t': 'NlbUnetConv',
    'nlb_unet_weights_t': 'NlbUnetWeights',
}


def _header_structs():
    """{struct name: [field names in declaration order]} parsed from the header's typedefs."""
    src = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    src = re.sub(r'//[^\n]*', '', src)
    out = {}
    for body, name in re.findall(r'typedef\s+struct\s*\{(.*?)\}\s*(nlb_\w+_t)\s*;', src, flags=re.S):
        fields = []
        for decl in body.split(';'):
            decl = decl.strip()
            if not decl:
                continue
            # "const float *a, *b" / "int N, S" / "float* base[2]" / "nlb_unet_conv_t inc[2]"
            first, *rest = decl.split(',')
            names = [re.sub(r'\[.*', '', first.split()[-1]).lstrip('*')] + [re.sub(r'\[.*', '', r.strip()).lstrip('* ') for r in rest]
            fields += [n for n in names if n]
        out[name] = fields
    return out


@pytest.mark.skipif(shutil.which('gcc') is None, reason='needs gcc')
def test_ctypes_mirrors_match_the_header(tmp_path):
    structs = _header_structs()
    assert set(MIRRORS) <= set(structs), sorted(set(MIRRORS) - set(structs))
    unmirrored = sorted(set(structs) - set(MIRRORS))
    assert not unmirrored, f'header structs without a checked ctypes mirror: {unmirrored}'
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "nlb200.h"', 'int main(void) {']
    for name, fields in structs.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for f in fields:
            lines.append(f'  printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run(['gcc', '-std=c99', '-I', os.path.dirname(HEADER), str(src), '-o', str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for name, cls_name in MIRRORS.items():
        cls = getattr(_lib, cls_name)
        assert C.sizeof(cls) == int(got[name]), (name, C.sizeof(cls), got[name])
        c_fields = structs[name]
        py_fields = [f[0] for f in cls._fields_]
        assert len(py_fields) == len(c_fields), (name, py_fields, c_fields)
        for pf, cf in zip(py_fields, c_fields):      # same order; offsets must agree field by field
            assert getattr(cls, pf).offset == int(got[f'{name}.{cf}']), (name, pf, cf)


def _header_prototypes():
    """{function: (return kind, [argument kinds])} with kinds 'ptr' / 'int' / 'u32' / 'i64' / 'size' / 'f32' / 'f64'."""
    src = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    src = re.sub(r'//[^\n]*', '', src)
    src = re.sub(r'typedef\s+struct\s*\{.*?\}\s*nlb_\w+_t\s*;', '', src, flags=re.S)

    def kind(t):
        t = t.strip()
        if '*' in t or '[' in t:
            return 'ptr'
        base = re.sub(r'\b(const|unsigned)\b', lambda m: m.group(0), t)
        words = base.replace('const', '').split()
        ty = ' '.join(words[:-1]) if len(words) > 1 else words[0]
        return {'int': 'int', 'int32_t': 'int', 'uint32_t': 'u32', 'unsigned int': 'u32', 'int64_t': 'i64', 'size_t': 'size',
                'float': 'f32', 'double': 'f64', 'void': 'void', 'char': 'char'}[ty]

    out = {}
    for ret, name, args in re.findall(r'([\w\s\*]+?)\s*\b(nlb_\w+)\s*\(([^;{]*?)\)\s*;', src, flags=re.S):
        args = ' '.join(args.split())
        arg_kinds = [] if args in ('', 'void') else [kind(a) for a in args.split(',')]
        out[name] = ('ptr' if '*' in ret else kind(ret + ' x'), arg_kinds)
    return out


def test_ctypes_signatures_match_the_header():
    """Argument count and kind (pointer / int / uint32 / int64 / float) of every entry of _lib.SIGNATURES against the
    prototypes of include/nlb200.h."""
    protos = _header_prototypes()
    assert set(_lib.SIGNATURES) <= set(protos), sorted(set(_lib.SIGNATURES) - set(protos))

    def ckind(t):
        if t is None:
            return 'void'
        if t in (C.c_void_p, C.c_char_p) or hasattr(t, 'contents') or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return 'ptr'
        return {C.c_int: 'int', C.c_int32: 'int', C.c_uint32: 'u32', C.c_int64: 'i64', C.c_size_t: 'size', C.c_float: 'f32',
                C.c_double: 'f64'}[t]

    for name, (res, args) in _lib.SIGNATURES.items():
        want_res, want_args = protos[name]
        got_args = [ckind(a) for a in args]
        assert got_args == want_args, (name, got_args, want_args)
        assert ckind(res) == want_res or (want_res == 'size' and ckind(res) in ('size', 'i64')), (name, ckind(res), want_res)
