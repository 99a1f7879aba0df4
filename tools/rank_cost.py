import os, sys
sys.path.insert(0, "/root/repo")
import torch, bench, phonic_b200
from phonic_b200 import workloads as W
from phonic_b200.player import Player
api = phonic_b200.load_api()
frames = W.frames_for(10, 48000)
out = torch.zeros(frames, 2, device="cuda")
for rank in (0, 0, 1, 1, 2, 3, 4, 5, 6, 7):
    p = Player(api, 48000); bench.build_scene(p, "cfg2", rank=rank, as_subtree=True)
    p.render_device(out.data_ptr(), frames); st = p.last_render_stats()
    print("rank", rank, "device %.2f skel %.2f replay %.2f fx %.2f" % (st.device_ms, st.skeleton_kernel_ms, st.voice_kernel_ms, st.effect_kernel_ms))
    p.close()
