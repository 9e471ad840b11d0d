import sys, ctypes, faulthandler
faulthandler.enable()
import sassy_b200
n = int(sys.argv[1])
use_torch = len(sys.argv) > 2 and sys.argv[2] == "torch"
s = sassy_b200.Searcher("dna", rc=False)
if use_torch:
    import torch
    host = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    host.fill_(65)
    addr = host.data_ptr()
else:
    addr = sassy_b200.host_alloc(n)
    ctypes.memset(addr, 65, n)
print("alloc ok", hex(addr), flush=True)
dt = s.upload_text((addr, n))
print("upload ok", flush=True)
print(len(s.search(b"ACGTACGTACGTACGTACGT", dt, 2)), s.stats(), flush=True)
print(len(s.search(b"ACGTACGTACGTACGTACGT", (addr, n), 2)), s.stats(), flush=True)
print(len(s.search(b"ACGTACGTACGTACGTACGT", (addr, n), 2)), flush=True)
print("done", flush=True)
