import numpy as np
W = np.load("gpurun_out/young_w_t8.npy").astype(np.float64)      # (100,100,784)
X = np.load("gpurun_out/young_x.npy").astype(np.float64)[:512]
gx, gy, D = W.shape
mu = np.full(D, 0.5)   # data mean of uniform [0,1)
def split16(a):
    amax = np.abs(a).max(axis=1, keepdims=True); amax[amax == 0] = 1
    sc = 2.0 ** (14 - np.floor(np.log2(amax)))
    hi = (a * sc).astype(np.float16).astype(np.float64); lo = (a * sc - hi).astype(np.float16).astype(np.float64)
    return hi / sc, lo / sc
xc = X - mu
xh, xl = split16(xc)
nxh, nxl = np.linalg.norm(xh, axis=1), np.linalg.norm(xl, axis=1)
def count(centres_per_neuron, label):
    Wf = W.reshape(-1, D)
    wc = Wf - centres_per_neuron                 # operand re-centred per neuron group
    wp = -2.0 * wc
    wh, wl = split16(wp)
    # exact score relative to global centre: |w-mu|^2 - 2 xc.(w-mu) = |w-mu|^2 - 2 xc.(w-c) - 2 xc.(c-mu)
    bias = ((Wf - mu) ** 2).sum(1)
    s_hat = xh @ wh.T + bias[None, :] - 2.0 * (xc @ (centres_per_neuron - mu).T)
    nwh, nwl = np.linalg.norm(wh, axis=1), np.linalg.norm(wl, axis=1)
    E = nxl[:, None] * nwh[None, :] + (nxh + nxl)[:, None] * nwl[None, :] + 2.0 ** -14 * nxh[:, None] * nwh[None, :]
    U = (s_hat + E).min(1, keepdims=True)
    cnt = ((s_hat - E) <= U).sum(1)
    print("%-44s candidates/row: mean %.1f median %d p90 %d max %d   mean |w - centre| %.4f" % (label, cnt.mean(), np.median(cnt), np.percentile(cnt, 90), cnt.max(), np.linalg.norm(wc, axis=1).mean()))
Wf = W.reshape(-1, D)
count(np.tile(mu, (gx * gy, 1)), "global centre (data mean)")
count(np.tile(Wf.mean(0), (gx * gy, 1)), "global centre (codebook mean)")
for pi, pj in ((10, 25), (5, 50), (16, 16), (4, 4), (2, 2)):
    C = np.zeros_like(W)
    for i0 in range(0, gx, pi):
        for j0 in range(0, gy, pj):
            C[i0:i0 + pi, j0:j0 + pj] = W[i0:i0 + pi, j0:j0 + pj].reshape(-1, D).mean(0)
    count(C.reshape(-1, D), "patch centres %dx%d" % (pi, pj))
