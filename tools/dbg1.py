import sys; sys.path.insert(0,'.')
import numpy as np, torch
from dsrg_b200 import api, synth
from oracle import crf_oracle
H,W,sf,img=41,41,12.0,'smooth'
B,M=3,21
batch = synth.make_batch(B,H,W,cues='cam',image=img,start=60)
probs = batch['probs'].copy(); probs[0,2,:3,:3]=1e-7
clamped = probs.copy(); clamped[clamped<1e-4]=1e-4
unary = np.transpose(clamped,(0,2,3,1)).copy()
want = np.stack([crf_oracle.CRF(batch['image'][b], unary[b], 10, sf) for b in range(B)])
eng = api.Engine(B,H,W,M)
d_im = torch.from_numpy(batch['image']).cuda()
for trial in range(3):
    d_un = torch.from_numpy(unary).cuda(); d_out = torch.empty_like(d_un)
    eng.crf_dev(d_un, d_im, api.crf_params(sf), d_out)
    got = d_out.cpu().numpy()
    d = np.abs(got-want)
    print('crf_dev NHWC trial',trial,'max',d.max(), 'per image',[float(d[b].max()) for b in range(B)], 'argmax', np.unravel_index(d.argmax(), d.shape))
d_p = torch.from_numpy(probs).cuda(); d_s = torch.empty_like(d_p); d_q = torch.empty_like(d_p)
eng.dsrg_forward_dev(torch.from_numpy(batch['labels']).cuda(), d_p, torch.from_numpy(batch['cues']).cuda(), d_im, api.crf_params(sf), 0.99,0.85,d_s,crf_out=d_q)
q = np.transpose(d_q.cpu().numpy(),(0,2,3,1))
d = np.abs(q-want)
print('dsrg_forward max',d.max(),[float(d[b].max()) for b in range(B)], np.unravel_index(d.argmax(), d.shape))
print('vs crf_dev', np.abs(q-got).max())
vs, vb = eng.lattice_sizes(B); print('V', vs, vb)
for b in range(B):
    c = crf_oracle.DenseCRF(W,H,M); c.set_unary_energy(-unary[b].ravel()); c.add_pairwise_energy(10,80/sf,80/sf,13,13,13,3,3/sf,3/sf,batch['image'][b].ravel())
    print(b, 'oracle V', c.lattice(0).M, c.lattice(1).M)
    for it in (1,2,3,5,10):
        qo = c.inference(it).reshape(H,W,M)
        d_un = torch.from_numpy(unary).cuda(); d_out = torch.empty_like(d_un)
        p = api.crf_params(sf, maxiter=it)
        eng.crf_dev(d_un, d_im, p, d_out)
        print('   iters',it,'max diff', float(np.abs(d_out.cpu().numpy()[b]-qo).max()))
