import sys
import numpy as np, torch
sys.path.insert(0, '.')
from tike_b200 import kernels as K, synthetic
from oracle import ptycho_np as onp
def dev(x): return torch.as_tensor(np.ascontiguousarray(x)).cuda()
for det, N, M in [(16,16,3)]:
    H, W = N + 40, N + 52
    psi, probe, scan = synthetic.make_problem(9, N, M, H, W, seed=det + M)
    far_ref = onp.farplane(psi, scan, probe, det)[:,0]
    b = K.make_batch(dev(psi[0]), dev(scan), dev(probe[0, 0]), det)
    far = torch.empty((9, M, det, det), dtype=torch.complex64, device='cuda')
    K.ptycho_fwd(b, far, None)
    f = far.cpu().numpy()
    # torch reference using the verified patch kernel
    patches = torch.zeros((9, N, N), dtype=torch.complex64, device='cuda')
    K.patch_fwd(dev(psi[0]), dev(scan), patches, N)
    ew = patches[:, None] * dev(probe[0, 0])[None]
    ft = torch.fft.fft2(ew, norm='ortho').cpu().numpy()
    for m in range(M):
        b1 = K.make_batch(dev(psi[0]), dev(scan), dev(probe[0, 0, m:m+1]), det)
        f1 = torch.empty((9, 1, det, det), dtype=torch.complex64, device='cuda')
        K.ptycho_fwd(b1, f1, None)
        f1 = f1.cpu().numpy()[:, 0]
        sc = np.abs(ft[:, m]).max()
        print('mode', m, 'gpu-multi vs torch', np.abs(f[:, m] - ft[:, m]).max() / sc,
              'gpu-single vs torch', np.abs(f1 - ft[:, m]).max() / sc,
              'oracle vs torch', np.abs(far_ref[:, m] - ft[:, m]).max() / sc)
