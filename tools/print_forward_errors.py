import os, sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from score_based_channels_b200 import params, entry_common as ec
from score_based_channels_b200.models import make_model
dev = torch.device('cuda:0')
G = 'tests/golden'
rel = lambda a, b: float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))
for name in ('forward_ngf8.npz', 'forward_ngf16.npz'):
    g = np.load(os.path.join(G, name))
    sd = params.random_state(int(g['ngf']), seed=int(g['wseed']))
    m = make_model(sd, ngf=int(g['ngf'])).to(dev)
    out = m(torch.from_numpy(g['x']).to(dev), torch.from_numpy(g['y']).to(dev)).cpu().numpy()
    print(name, ['%.2e' % rel(out[b], g['out'][b]) for b in range(out.shape[0])])
g = np.load(os.path.join(G, 'real_ckpt_forward.npz'))
c = ec.load_checkpoint('fixtures_local/score-deepest-cdl-c.pt')
m = ec.build_model(c['config'], c['model_state'], dev, 'tf32x3')
out = m(torch.from_numpy(g['x']).to(dev), torch.from_numpy(g['y']).to(dev)).cpu().numpy().astype(np.float64)
print('real ckpt: err/e32', ['%.2f' % (rel(out[b], g['out64'][b]) / rel(g['out32'][b].astype(np.float64), g['out64'][b])) for b in range(6)])
