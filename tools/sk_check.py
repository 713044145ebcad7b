import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import compute_war_b200 as cw, oracle_lib as O
e = cw.Engine(0)
rng = np.random.default_rng(0)
bad=[]
for n in list(range(1,200))+[4096,8192]:
    d = rng.integers(0,256,n*3,dtype=np.uint8)
    g = e.do_hashing(cw.HASH_SKEIN512, d, n); o = O.hash_chunks(O.HASH_SKEIN512, d, n)
    g2 = e.do_hashing(cw.HASH_SKEIN256_128, d, n); o2 = O.hash_chunks(O.HASH_SKEIN256_128, d, n)
    if not (g==o).all() or not (g2==o2).all(): bad.append((n, (g==o).all(axis=1).tolist(), (g2==o2).all(axis=1).tolist()))
print("bad:", bad[:20], len(bad))
