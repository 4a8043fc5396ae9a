import os, time, sys, numpy as np
sys.path.insert(0,'.')
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ["/sys/fs/cgroup/cpu.max","/sys/fs/cgroup/cpu/cpu.cfs_quota_us","/sys/fs/cgroup/cpuset.cpus.effective"]:
    try: print(f, open(f).read().strip())
    except Exception as e: print(f, "n/a")
os.system("lscpu | egrep 'Model name|Socket|Core|Thread|NUMA|MHz' | head -12; cat /proc/loadavg; free -g | head -2")
from oracle.bindings import RefLib
from scalablevectorsearch_b200.synthetic import clustered_unit_vectors
ref=RefLib()
base,queries=clustered_unit_vectors(100000,10000,96)
for th in [8,16,32,64]:
    t=time.time(); g,ep=ref.build(base,'l2',64,128,alpha=1.2,threads=th); print("build 100k threads",th, round(time.time()-t,1),"s", flush=True)
for th in [1,4,8,16,32,64,128]:
    idx=ref.index(base,g,ep,'l2',threads=th)
    idx.search(queries,10,128,128)
    ts=[]
    for _ in range(3):
        t=time.perf_counter(); idx.search(queries,10,128,128); ts.append(time.perf_counter()-t)
    print("search threads",th, round(10000/min(ts)), "QPS", flush=True)
