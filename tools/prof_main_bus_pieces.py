"""Main-bus stage fed piece by piece (as render_sharded does): wall time per piece against its device span (debug aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import phonic_b200
from phonic_b200 import workloads as W
from phonic_b200.distributed import MainBusStage, piece_bounds
api = phonic_b200.load_api()
frames = W.frames_for(10, 48000)
pf = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
bus = (torch.randn(frames, 2, device="cuda") * 0.1).contiguous(); out = torch.zeros_like(bus)
for it in range(3):
    st = MainBusStage(api, 48000, W.add_main_bus_sends)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for off, n in piece_bounds(frames, pf):
        st.process(bus[off:off + n], out[off:off + n])
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) * 1e3
    print("pieces of %d: wall %.1f ms, device %.1f ms" % (pf, dt, st.device_ms))
    st.close()
