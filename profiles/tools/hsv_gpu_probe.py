"""GPU probe: K1's HueSaturationValue on all 2^24 RGB triples (identity resize of a 4096 x 4096 image) against the same
pixel function compiled for the host (which tests/test_transforms.py pins to cv2)."""
import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from nkb_classification_b200 import ops, transforms as T
dev=torch.device('cuda:0')
g=np.arange(256,dtype=np.uint8)
rgb=np.stack(np.meshgrid(g,g,g,indexing='ij'),-1).reshape(1,4096,4096,3)
plan=T.compile_pipeline([T.Resize(4096,4096), T.HueSaturationValue(hue_shift_limit=20, sat_shift_limit=10, val_shift_limit=50, p=1.0), T.Normalize(), T.ToTensorV2()])
import random
b=plan.draw(1, random.Random(3))
print('flags',b.flags, b.hsv_shift)
u8=torch.zeros((1,4096,4096,3),dtype=torch.uint8,device=dev)
out=ops.preprocess_crops(torch.from_numpy(rgb).to(dev), torch.tensor([[0,0,4096,4096]],dtype=torch.int32,device=dev), torch.zeros(1,dtype=torch.int32,device=dev), plan, out_u8=u8, aug=b)
torch.cuda.synchronize()
got=u8.cpu().numpy().reshape(-1,3)
exp=ops.debug_hsv_shift(rgb.reshape(-1,3), b.hsv_lut[0], b.hsv_trunc_cols > 0)
mis=(got!=exp).any(1)
print('mismatches', int(mis.sum()))
idx=np.nonzero(mis)[0][:10]
for i in idx: print(rgb.reshape(-1,3)[i], got[i], exp[i])
