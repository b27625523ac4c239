import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch, ctypes as C
import zebrapose_b200 as zp
sys.argv=['x']; import bench
logits,bboxes,Ks,obj,tables,crops=bench.make_workload(64,1002)
eng=zp.Engine(0)
for j,t in enumerate(tables): eng.upload_dict(j,t)
lg=torch.from_numpy(logits).cuda()
corr,counts=eng.decode(lg,bboxes,obj.astype(np.int32))
for _ in range(3): r=eng.ransac(corr,counts,Ks.reshape(-1,9))
buf=(C.c_int64*24)()
eng.lib.zp_debug_clocks(eng.ctx.handle, buf)
c=np.array(list(buf))
names=['select','pass0','pass1','pca','pass2','jacobi','pick4','cands','pass3+pick']
d=np.diff(c[:10])
for n,v in zip(names,d): print('%-10s %8d cycles %7.1f us'%(n,v,v/1965.0))
print('total', (c[9]-c[0])/1965.0,'us')
nm=['setup','nullspace(jacobi)','L_rho','candidates','errors+write']
for n,v in zip(nm,np.diff(c[10:16])): print('minimal %-18s %8d cycles %7.1f us'%(n,v,v/1965.0))
nm=['tridiag','bisection','invit+backtransform']
for n,v in zip(nm,np.diff(c[16:20])): print('eig in final (16 lanes) %-20s %8d cycles %7.1f us'%(n,v,v/1965.0))
cap=corr.shape[2]
smp=eng.make_samples(counts,cap,150,5)
for _ in range(2): eng.solve_minimal(corr,counts,torch.from_numpy(Ks.reshape(-1,9)).cuda(),smp)
eng.lib.zp_debug_clocks(eng.ctx.handle, buf)
c=np.array(list(buf))
for n,v in zip(nm,np.diff(c[16:20])): print('eig in minimal (quads)  %-20s %8d cycles %7.1f us'%(n,v,v/1965.0))
