"""Times dfol_pair_hidden_fwd_tc / dfol_pair_hidden_bwd_tc alone at the bench shapes (B200 only).

    python tools/time_pair_hidden.py [n_objects] [images]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dfol_vqa_b200.capi import call, ptr, stream_ptr  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    H = 256
    T, P = n * B, n * n * B
    dev = 'cuda'
    uv = torch.randn(T, 2 * H, device=dev) * 0.5
    obj = torch.zeros(T, 8, device=dev)
    obj[:, 4:] = torch.rand(T, 4, device=dev)
    w = torch.randn(H, 12, device=dev) * 0.3
    bias = torch.randn(H, device=dev) * 0.1
    h = torch.empty(P, H, device=dev, dtype=torch.bfloat16)
    geo = torch.empty(P, 4, device=dev)
    cnt = torch.full((B,), n, dtype=torch.long)
    img_n = cnt.to(torch.int32).cuda()
    obj_row = torch.cat([torch.zeros(1, dtype=torch.long), cnt.cumsum(0)]).to(torch.int32).cuda()
    pair_row = torch.cat([torch.zeros(1, dtype=torch.long), (cnt * cnt).cumsum(0)]).to(torch.int32).cuda()
    dz = (torch.randn(P, H, device=dev) * 0.1).bfloat16()
    dcat = torch.empty(T, 2 * H, device=dev, dtype=torch.bfloat16)
    dwg = torch.zeros(H, 12, device=dev)
    db = torch.zeros(H, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def fwd():
        call('dfol_pair_hidden_fwd_tc', ptr(uv), 2 * H, ptr(obj[:, 4:]), 8, ptr(w[:, 8:]), 12, ptr(bias), ptr(h), H, H,
             ptr(geo), ptr(pair_row), ptr(obj_row), ptr(img_n), B, n, stream_ptr())

    def bwd():
        call('dfol_pair_hidden_bwd_tc', ptr(dz), H, ptr(geo), ptr(dcat), ptr(dcat[:, H:]), 2 * H, ptr(dwg[:, 8:]), 12,
             ptr(db), H, ptr(pair_row), ptr(obj_row), ptr(img_n), B, n, stream_ptr())

    for name, fn, nbytes in (('fwd', fwd, P * H * 2 + P * 16), ('bwd', bwd, P * H * 2 + P * 16)):
        ts = []
        for it in range(13):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        med = ts[len(ts) // 2]
        print('%s n=%d B=%d  %.4f ms  %.0f GB/s  (DFOL_PF_SUBJECTS=%s)' % (name, n, B, med, nbytes / med / 1e6,
                                                                           os.environ.get('DFOL_PF_SUBJECTS', '-')))


if __name__ == '__main__':
    main()
