"""Large-N check of the tensor-core search against the exact search + fallback-row count."""
import torch, sys, os
sys.path.insert(0, '/root/repo')
from favae_b200 import _lib
torch.manual_seed(0)
K, D = 16384, 256
for n in (8192, 65536, 262144):
    x = torch.randn(n, D, device='cuda'); e = torch.nn.functional.normalize(torch.randn(K, D, device='cuda'), dim=-1)
    xn = torch.empty(n, D, device='cuda'); xh = torch.empty(n, D, device='cuda', dtype=torch.float16)
    en = torch.empty(K, D, device='cuda'); eh = torch.empty(K, D, device='cuda', dtype=torch.float16)
    st = _lib.stream()
    _lib.call('favae_vq_prepare_rows', x.data_ptr(), n, D, 1, 1, xn.data_ptr(), xh.data_ptr(), None, st)
    _lib.call('favae_vq_prepare_rows', e.data_ptr(), K, D, 1, 1, en.data_ptr(), eh.data_ptr(), None, st)
    nb = _lib.load().favae_vq_search_tc_workspace_bytes(n, K, D)
    ws = torch.full((nb,), 0xAB, device='cuda', dtype=torch.uint8)
    idx = torch.empty(n, device='cuda', dtype=torch.int64); keys = torch.empty(n, device='cuda', dtype=torch.int64)
    _lib.call('favae_vq_search_tc', xh.data_ptr(), eh.data_ptr(), xn.data_ptr(), en.data_ptr(), n, K, D, ws.data_ptr(), nb, keys.data_ptr(), idx.data_ptr(), st)
    torch.cuda.synchronize()
    # ovf_count lives at the last aligned block: find by scanning tail ints
    idx2 = torch.empty_like(idx)
    _lib.call('favae_vq_search_exact', xn.data_ptr(), en.data_ptr(), None, n, K, D, 0, keys.data_ptr(), idx2.data_ptr(), st)
    torch.cuda.synchronize()
    import ctypes
    cnt = ctypes.c_int(-1)
    _lib.call('favae_vq_search_tc_overflow_rows', ws.data_ptr(), n, K, D, ctypes.addressof(cnt))
    print(n, 'mismatch', int((idx != idx2).sum()), 'fallback rows', cnt.value)
