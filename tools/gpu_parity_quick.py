import sys, time, ctypes as C
sys.path.insert(0, 'multimodal-scene-captioning_b200'); sys.path.insert(0, '.')
import numpy as np, torch
from msc_geom.synthetic import make_sample
from msc_geom.layout import pack_batch, GeomParams
from msc_geom.engine import GeometryEngine
from msc_geom import _capi
eng = GeometryEngine()
import os
_capi.set_option('config', int(os.environ.get('MSC_CONFIG','9')))
_capi.set_option('split', int(os.environ.get('MSC_SPLIT','0')))
print('device', torch.cuda.get_device_name(), 'sms', eng.sm_count, 'smem', eng.smem_optin)
samples = [make_sample(i, n_sweeps=10 if i%2==0 else 3, n_boxes='mini' if i%3==0 else 60) for i in range(5)]
hb = pack_batch(samples)
p = GeomParams()
t=time.time(); out = eng.process_samples(samples); print('gpu s', time.time()-t, 'window', _capi.get_option('last_window'), 'smem', _capi.get_option('last_smem'), 'split', _capi.get_option('last_split'), 'config', _capi.get_option('last_config'))
from tests.oracle_bridge import oracle_fused
ok=True
for i,s in enumerate(samples):
    ref = oracle_fused(hb, i, p)
    b0,b1 = hb.sample_box_off[i], hb.sample_box_off[i+1]
    for k in ('box_count','box_nearest','box_centroid','proj_visible','proj_extent'):
        a=out[k][b0:b1]; r=ref[k]
        same = np.array_equal(a, r)
        print(i,k,'EXACT' if same else 'DIFF', '' if same else (np.abs(a.astype(np.float64)-r.astype(np.float64)).max()))
        ok&=same
    for k in ('bev_count','bev_isum_q','bev_height'):
        a=out[k][i]; r=ref[k]; same=np.array_equal(a,r); print(i,k,'EXACT' if same else 'DIFF %d cells'%(a!=r).sum()); ok&=same
    a=out['stats'][i][:14]; r=ref['stats'][:14]; same=np.array_equal(a,r); print(i,'stats',a[:13],'EXACT' if same else ('DIFF',r[:13])); ok&=same
print('ALL OK' if ok else 'MISMATCH')
