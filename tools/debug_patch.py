import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))
from b200timm import ops
torch.manual_seed(0)
B, C, H, W, p = 2, 3, 224, 224, 16
K = C * p * p
# image whose value encodes (c, y, x) position: small integers exactly representable in bf16
c = torch.arange(C).view(1, C, 1, 1).float()
y = torch.arange(H).view(1, 1, H, 1).float()
x = torch.arange(W).view(1, 1, 1, W).float()
img = ((y % 16) * 16 + (x % 16)).expand(B, C, H, W).contiguous()      # = ky*16+kx in [0,255] exact in bf16
img = img.cuda().bfloat16()
Wm = torch.eye(K, device="cuda").bfloat16()                            # D = K = 768: out row = the A row itself
out = ops.patch_embed_fwd(img, Wm, None, None, 0, 196, 0, p)
ref = ops.patchify(img, p)                                             # [B*196, 768]
ref = ref.view(B, 196, K)
print("max err", (out.float() - ref.float()).abs().max().item())
o = out[0].float().cpu(); r = ref[0].float().cpu()
for g in (0, 1, 13, 14, 97, 98, 195):
    print("row", g, "got", o[g, :20].tolist(), "\n      want", r[g, :20].tolist())
# which source k does each output k hold (values encode ky*16+kx for channel 0 block)
row = o[0, :256]
print("row0 first 64:", row[:64].tolist())
img2 = torch.randn(B, C, H, W, device="cuda").bfloat16()
out2 = ops.patch_embed_fwd(img2, Wm, None, None, 0, 196, 0, p); ref2 = ops.patchify(img2, p).view(B, 196, K)
bad = (out2.float() - ref2.float()).abs() > 1e-6
print("random image: mismatching fraction", bad.float().mean().item(), "rows with any mismatch", bad.any(-1).float().mean().item())
print("mismatch by k-block:", bad.view(B, 196, 12, 64).any(-1).float().mean((0, 1)).tolist())
print("mismatch by row mod 8:", [bad.view(B * 196, K)[i::8].float().mean().item() for i in range(8)])
