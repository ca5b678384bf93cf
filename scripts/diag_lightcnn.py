import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/golden')
import numpy as np, torch
import model_cases as MC
from ffwm_b200 import light_cnn as L, conv
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
gold = np.load('/root/repo/tests/golden/ref_models_f64.npz')
for en in (True, False):
    conv.ENABLED = en
    got = MC.run_lightcnn(MC.fill_state(L.LightCNN_29Layers(num_classes=100), torch.float32).to('cuda:0'))
    for k in ('fc', 'pool', 'grad/x'):
        w = gold['lightcnn/' + k]; g = got[k]
        print('tcgen05' if en else 'cudnn  ', k, 'max-rel %.3e' % (np.abs(g - w).max() / np.abs(w).max()), 'l2-rel %.3e' % (np.linalg.norm(g - w) / np.linalg.norm(w)))
