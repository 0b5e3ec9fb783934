import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neural-astar_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch, torch.nn.functional as F
torch.manual_seed(0)
x = torch.randn(100, 256, 32, 32, device="cuda").contiguous(memory_format=torch.channels_last).relu_()
w = torch.randn(1, 256, 3, 3, device="cuda") * 0.05
b = torch.randn(1, device="cuda")
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(n):
        a,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e)*1e3)
    return float(np.median(ts))
sel = torch.zeros(1, 9, 3, 3, device="cuda")
for k in range(9): sel[0, k, k // 3, k % 3] = 1
Wm = w[0].reshape(256, 9).contiguous()
def gemm_path():
    B, C, H, W_ = x.shape
    t2 = x.permute(0, 2, 3, 1).reshape(-1, C) @ Wm              # [B*H*W, 9]
    tp = t2.view(B, H, W_, 9).permute(0, 3, 1, 2)               # [B, 9, H, W] (channels-last strides)
    with torch.backends.cudnn.flags(enabled=True, benchmark=True, deterministic=False, allow_tf32=False):
        return F.conv2d(tp, sel, b, padding=1)
def gemm_path_adds():
    B, C, H, W_ = x.shape
    t2 = (x.permute(0, 2, 3, 1).reshape(-1, C) @ Wm).view(B, H, W_, 9)
    tp = F.pad(t2, (0, 0, 1, 1, 1, 1))
    out = b.view(1, 1, 1).expand(B, H, W_).clone()
    for k in range(9):
        ky, kx = divmod(k, 3)
        out = out + tp[:, ky:ky + H, kx:kx + W_, k]
    return out.unsqueeze(1)
ref = F.conv2d(x, w, b, padding=1)
with torch.backends.cudnn.flags(enabled=True, benchmark=False, deterministic=False, allow_tf32=False):
    ref32 = F.conv2d(x, w, b, padding=1)
g1 = gemm_path(); g2 = gemm_path_adds()
print("max|ref(tf32)-ref32|", float((ref - ref32).abs().max()), " max|gemm-ref32|", float((g1 - ref32).abs().max()), " adds:", float((g2 - ref32).abs().max()), "scale", float(ref32.abs().max()))
print("cudnn conv2d (tf32):", t(lambda: F.conv2d(x, w, b, padding=1)))
print("gemm + select-conv :", t(gemm_path))
print("gemm + 9 adds      :", t(gemm_path_adds))
print("gemm only          :", t(lambda: x.permute(0, 2, 3, 1).reshape(-1, 256) @ Wm))
torch.backends.cuda.matmul.allow_tf32 = True
print("gemm only (tf32 matmul):", t(lambda: x.permute(0, 2, 3, 1).reshape(-1, 256) @ Wm))
