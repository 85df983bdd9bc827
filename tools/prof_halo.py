"""Halo exchange over NVLink in ONE process (so that ncu can watch it): two slabs of the 3-D 7-pt
Laplacian n^3 (cfg4's wire size per neighbour: n^2 doubles = 2 MiB at n = 512), subdomain 0 on
GPU 0, subdomain 1 on GPU 1 (both on GPU 0 when the box has one).  Times the pack + peer-store
push, the unpack and the pair with CUDA events; under ncu the same launches carry the NVLink byte
counters:

    python tools/prof_halo.py [n=512] [reps=20]
    ncu --metrics gpu__time_duration.sum,nvltx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes.sum \
        --clock-control none -k regex:halo_ --csv --log-file gpurun_out/halo_nvlink.csv \
        python tools/prof_halo.py 512 5
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "schwarz-lib_b200"))
import schwz_b200 as S

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
P = 2
ndev = min(S.device_count(), 2)
setup = S.Setup(("laplacian3d", n), P)
ctxs = [S.Context(r % ndev) for r in range(P)]
if ndev > 1:
    S.enable_peers(ctxs)
subs = []
for r in range(P):
    subs.append(S.Ras(ctxs[r], setup, r, local_max_iters=5))
    setup.release(r)
S.connect_local(subs, setup)
# one real exchange first (epochs, peer tables)
S.ras_run(subs, P, 2, tolerance=1e-30, enable_global_check=True)
s = subs[0]
payload = 8 * len(setup.put_list(0, 0))
print("devices %d, payload per push %d bytes (%.2f MiB)" % (ndev, payload, payload / 2 ** 20))
for kind, name in ((5, "pack + push (peer stores) + publish"), (6, "unpack"), (4, "push + unpack")):
    ms = s.kernel_time_ms(kind, reps)
    print("%-36s %8.2f us   payload %7.1f GB/s%s"
          % (name, ms * 1e3, payload / ms / 1e6,
             "   = %.2f of 900 GB/s NVLink, %.2f of the 770 GB/s measured peer copy"
             % (payload / ms / 1e6 / 900, payload / ms / 1e6 / 770) if kind == 5 and ndev > 1 else ""))
for x in subs:
    x.close()
for c in ctxs:
    c.close()
