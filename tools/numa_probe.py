"""NUMA / PCIe probe for the GPU box: which node each GPU hangs off, whether set_mempolicy/mbind are permitted in the
container, and what node-local vs node-remote pinned staging costs in H2D/D2H bandwidth (one GPU, through the C ABI's
own allocator and copy entry points).  Output: JSON on stdout.  python tools/numa_probe.py"""
import ctypes
import glob
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
libc = ctypes.CDLL(None, use_errno=True)
SYS_mbind, SYS_set_mempolicy, SYS_get_mempolicy = 237, 238, 239
MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND = 0, 1, 2


def set_mempolicy(mode, node):
    mask = ctypes.c_ulong(0 if node is None else 1 << node)
    r = libc.syscall(SYS_set_mempolicy, mode, ctypes.byref(mask) if node is not None else None, 64 if node is not None else 0)
    return r, ctypes.get_errno()


def node_of_addr(addr):
    mode = ctypes.c_int(0)
    MPOL_F_NODE, MPOL_F_ADDR = 1, 2
    r = libc.syscall(SYS_get_mempolicy, ctypes.byref(mode), None, 0, ctypes.c_void_p(addr), MPOL_F_NODE | MPOL_F_ADDR)
    return mode.value if r == 0 else -1


out = {}
out["nodes"] = sorted(os.path.basename(p) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
out["node_cpus"] = {os.path.basename(p): open(p + "/cpulist").read().strip() for p in glob.glob("/sys/devices/system/node/node[0-9]*")}
out["affinity"] = len(os.sched_getaffinity(0))
out["cpu_count"] = os.cpu_count()
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpuset.cpus.effective", "/sys/fs/cgroup/cpuset.mems.effective"):
    try:
        out[f] = open(f).read().strip()
    except Exception as e:
        out[f] = "n/a: %s" % e
try:
    q = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout
    gp = {}
    for line in q.strip().splitlines():
        idx, bdf = [x.strip() for x in line.split(",")]
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]
        try:
            gp[idx] = {"bdf": bdf, "numa_node": open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip()}
        except Exception as e:
            gp[idx] = {"bdf": bdf, "numa_node": "n/a: %s" % e}
    out["gpus"] = gp
except Exception as e:
    out["gpus"] = "n/a: %s" % e
out["meminfo"] = {os.path.basename(os.path.dirname(p)): open(p).read().split("\n")[0:2] for p in glob.glob("/sys/devices/system/node/node*/meminfo")}

import spark_s3_shuffle_b200 as pkg

c = pkg.capi
c.init(1)
L = c.load()
N = 4 << 30
d = c.dev_alloc(N)
res = {}
for node in [None] + list(range(len(out["nodes"]))):
    if node is not None:
        r, e = set_mempolicy(MPOL_BIND, node)
        out["set_mempolicy_bind_%d" % node] = [r, e]
        if r != 0:
            continue
    p = L.b2s_host_alloc(N)
    where = [node_of_addr(p + k * (N // 4)) for k in range(4)]
    set_mempolicy(MPOL_DEFAULT, None)
    bw = {}
    for kind, name in ((1, "h2d"), (2, "d2h")):
        best = 0
        for _ in range(3):
            t = time.perf_counter()
            if kind == 1:
                c.dev_memcpy(d, p, N, 1)
            else:
                c.dev_memcpy(p, d, N, 2)
            best = max(best, N / (time.perf_counter() - t) / 1e9)
        bw[name] = round(best, 2)
    res[str(node)] = {"pages_on_node": where, "GBps": bw}
    L.b2s_host_free(p)
out["pinned_copy_by_policy"] = res
print(json.dumps(out, indent=1))
