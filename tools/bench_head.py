"""Times the pieces of the MVDualAttAlignment offset/mask head at c3, B = 12."""
import ctypes, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cdfo_b200 import _lib, conv  # noqa: E402
from tools.bench_conv import timeit  # noqa: E402

dev = torch.device("cuda:0")
B, H, W, dg = 12, 272, 480, 16
o = torch.randn(2 * B, 64, H, W, device=dev)
w0 = torch.randn(64, 64, 3, 3, device=dev) * 0.05
w2 = torch.randn(432, 64, 3, 3, device=dev) * 0.02
b0, b2 = torch.randn(64, device=dev), torch.randn(432, device=dev)
res = {}
res["to_c8_2B"] = timeit(lambda: conv.to_c8(o))
o8 = conv.to_c8(o)
res["conv0_lrelu_2B"] = timeit(lambda: conv.conv3x3(o8, w0, b0, conv.ACT_LRELU))
z = conv.conv3x3(o8, w0, b0, conv.ACT_LRELU)
wpk = conv.pack_weight(w2)
first = torch.empty((B, 9, 8, H, W, 2, 4), device=dev, dtype=torch.float16)
fields = torch.empty_like(first)
args = (B, 64, dg, H, W, ctypes.c_float(10.0), _lib.stream_ptr(dev))
res["head1"] = timeit(lambda: _lib.call("cdfo_mv_offset_head_sm100_fwd", _lib.ptr(z[:B]), _lib.ptr(wpk), _lib.ptr(b2), _lib.ptr(None), _lib.ptr(first), *args))
res["head2"] = timeit(lambda: _lib.call("cdfo_mv_offset_head_sm100_fwd", _lib.ptr(z[B:]), _lib.ptr(wpk), _lib.ptr(b2), _lib.ptr(first), _lib.ptr(fields), *args))
res["conv2_plain_c8out"] = timeit(lambda: conv.conv3x3(z[:B], w2, b2, 0))
res["conv2_plain_nchw_out"] = timeit(lambda: conv.conv3x3(z[:B], w2, b2, 0, out_nchw=True))
res["cat"] = timeit(lambda: torch.cat([o[:B], o[B:]], 0))
print(json.dumps({k: round(v, 1) for k, v in res.items()}))
