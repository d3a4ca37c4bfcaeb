"""Dev tool (GPU): render throughput of BASELINE configs[2] (LiDAR-only render of a
32 x 1084 nuScenes sweep: depth + intensity + semantics) and of one 1600 x 900 camera
frame (configs[4] is four of these + a sweep), through models.render_image.
  python tools/render_bench.py [chunk_size]"""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_lidar_b200 import configs, models, synthetic

cfg = configs.nuscenes_single()
if len(sys.argv) > 1:
    cfg.render_chunk_size = int(sys.argv[1])
model = models.Model(cfg).cuda()
model.load_state_dict({k: v.cuda() for k, v in synthetic.init_state_dict(seed=0, table_std=0.05).items()}, strict=False)
model.graph_render = os.environ.get('NLB_RENDER_GRAPH', '1') != '0'   # 0: eager launches per chunk


def timeit(fn, n):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out

res = {}
sweep = {k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_lidar_sweep(seed=0)).items()}
n = sweep['origins'].shape[0]
ms, out = timeit(lambda: models.render_image(model, None, sweep, False, cfg, image=False, verbose=False), 10)
res['lidar_sweep'] = dict(rays=n, ms=ms, rays_per_s=n / ms * 1e3, chunk=cfg.render_chunk_size,
                          outputs=sorted(k for k in out if not k.startswith('ray_')))
frame = {k: v.cuda() for k, v in synthetic.to_torch(synthetic.make_camera_frame(seed=0)).items()}
nf = frame['origins'].shape[0]
ms, out = timeit(lambda: models.render_image(model, None, frame, False, cfg, image=False, verbose=False), 3)
res['camera_frame_1600x900'] = dict(rays=nf, ms=ms, rays_per_s=nf / ms * 1e3, chunk=cfg.render_chunk_size)
print(json.dumps(res))
