import os, sys
sys.path.insert(0, '/root/repo')
os.environ["RNB_FUSE"]=sys.argv[1]; os.environ["RNB_FUSE_NEXT"]=sys.argv[2]; os.environ["RNB_AUTOTUNE"]="0"
import torch
from resnet_c_b200 import engine, weights
b=int(sys.argv[3])
m = engine.ResNet("resnet50", weights.cached_weights_dir("resnet50", 0, True), dtype="bf16", max_batch=b)
x = weights.synthetic_images(b).cuda()
l,t = m.forward(x); torch.cuda.synchronize(); print(t[:8])
