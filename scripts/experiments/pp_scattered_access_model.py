"""CPU model of the PP history pass's scattered accesses (no GPU needed): for every warp of 32
history points, how many distinct 32-byte sectors one table-read instruction and one round of
32 candidate loads touch -- in the order the points arrive (beam-major), after sorting chunks of
one traversal by cell, and after a merged sort of all 16 traversals of the scan.  The beam-order
row reproduces what ncu measures on the shipped kernel (18.8 sectors per column-record load,
22 per dealt candidate round: profiles/r1b_pp_count_variants_ncu_full_summary.csv, source page).
Usage: python scripts/experiments/pp_scattered_access_model.py      (a few minutes of NumPy)"""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
from modest_b200 import synth
case = synth.make_scan_case(1000, synth.LYFT, n_traversals=4, frames_per_traversal=1, n_points=60000)
q = case.query_fixed.astype(np.float32)
G=512; cell=np.float32(0.3*1.001); inv=np.float32(1)/cell
lo=q.min(0); hi=q.max(0)
x0=np.float32(0.5)*(lo[0]+hi[0])-np.float32(0.5)*cell*G; y0=np.float32(0.5)*(lo[1]+hi[1])-np.float32(0.5)*cell*G; z0=lo[2]
def cells(p):
    cx=np.clip(np.floor((p[:,0]-x0)*inv).astype(np.int64),1,G-2)
    cy=np.clip(np.floor((p[:,1]-y0)*inv).astype(np.int64),1,G-2)
    cz=np.clip(np.floor((p[:,2]-z0)*inv).astype(np.int64),0,31)
    return cx,cy,cz
qcx,qcy,qcz=cells(q)
# sorted query order (y,x,z)
key=(qcy*G+qcx)*32+qcz
order=np.argsort(key,kind='stable'); skey=key[order]
# cell start lookup via searchsorted
def cand_ranges(hcx,hcy,hcz):
    # for each history point: list of (start,end) per column (9), z window cz-1..cz+1
    out=[]
    for dy in (-1,0,1):
        for dx in (-1,0,1):
            col=(hcy+dy)*G+(hcx+dx)
            za=np.clip(hcz-1,0,31); zb=np.clip(hcz+1,0,31)
            s=np.searchsorted(skey, col*32+za, 'left'); e=np.searchsorted(skey, col*32+zb, 'right')
            out.append((s,e))
    return out
h=case.history[1].astype(np.float32)
def analyse(h, name):
    hcx,hcy,hcz=cells(h)
    n=len(h)//32*32
    W=n//32
    # record loads: 9 instructions, address = ((hcy+dy)*G + hcx+dx)*8 bytes
    rec_sect=[]; 
    for dy in (-1,0,1):
        for dx in (-1,0,1):
            addr=((hcy[:n]+dy)*G+(hcx[:n]+dx))*8
            sec=(addr//32).reshape(W,32)
            rec_sect.append(np.array([len(np.unique(r)) for r in sec]))
    rec=np.mean(rec_sect)
    # candidates: per point list of candidate positions; warp-level: deal evenly -> count distinct sectors per round of 32
    rs=cand_ranges(hcx[:n],hcy[:n],hcz[:n])
    tot_sect=0; tot_rounds=0; tot_c=0
    for w in range(0,W,7):   # sample warps
        pos=[]
        for i in range(w*32,(w+1)*32):
            for s,e in rs:
                if e[i]>s[i]: pos.extend(range(s[i],e[i]))
        pos=np.array(pos)
        if len(pos)==0: continue
        T=len(pos); S=(T+31)//32
        # dealing: lane l takes slice [l*S,(l+1)*S): round r = elements l*S + r
        for r in range(S):
            idx=np.arange(32)*S+r; idx=idx[idx<T]
            tot_sect+=len(np.unique(pos[idx]*16//32)); tot_rounds+=1
        # flat: consecutive
        tot_c+=T
    # flat-list mode
    flat_sect=0; flat_rounds=0
    for w in range(0,W,7):
        pos=[]
        for i in range(w*32,(w+1)*32):
            for s,e in rs:
                if e[i]>s[i]: pos.extend(range(s[i],e[i]))
        pos=np.array(pos)
        for r in range(0,len(pos),32):
            flat_sect+=len(np.unique(pos[r:r+32]*16//32)); flat_rounds+=1
    print(f"{name:28s} record-load sectors/instr {rec:5.1f} | dealt candidates: {tot_sect/max(tot_rounds,1):5.1f} sectors/round, {tot_rounds/len(range(0,W,7)):4.1f} rounds/warp | flat list: {flat_sect/max(flat_rounds,1):5.1f} sectors/round, {flat_rounds/len(range(0,W,7)):4.1f} rounds/warp")
analyse(h, "beam order (as given)")
hcx,hcy,hcz=cells(h)
hk=(hcy*G+hcx)*32+hcz
for chunk in (256,1024,4096,16384,len(h)):
    hh=h.copy()
    for c0 in range(0,len(h),chunk):
        sl=slice(c0,min(len(h),c0+chunk))
        o=np.argsort(hk[sl],kind='stable'); hh[sl]=h[sl][o]
    analyse(hh, f"sorted by cell in chunks of {chunk}")
print("--- merged sort over all traversals of the scan (16 traversals) ---")
case16 = synth.make_scan_case(1000, synth.LYFT, n_traversals=16, frames_per_traversal=1, n_points=60000)
H = np.concatenate(case16.history).astype(np.float32)
hcx,hcy,hcz=cells(H); hk=(hcy*G+hcx)*32+hcz
o=np.argsort(hk,kind='stable')
Hs=H[o]
# analyse only a slice to keep the python loops short
analyse(Hs[:120000], "merged 16-traversal sort")
