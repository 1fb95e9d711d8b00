import sys, time; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/ml-4m_b200')
import bench
for th, B in ((16,8),(32,8),(64,8),(32,16)):
    t=time.time(); v,sec,c = bench.cpu_reference_steps(1,1,B,128,threads=th); print(th,B,'tok/s',round(v,1),'sec/step',round(sec,2),'total',round(time.time()-t,1), flush=True)
