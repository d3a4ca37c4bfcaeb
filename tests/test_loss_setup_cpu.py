"""Host logic of the loss assembly (no GPU): which entries of [data, depth, sem, int, d_smo, s_smo] a configuration
uses and the step-dependent multipliers of Z/train.py:330-371 that the fused supervision kernel receives."""
import torch

from nerf_lidar_b200 import configs, train


def _rend(sem=True, inten=True):
    r = dict(rgb=torch.zeros(4, 3), depth=torch.zeros(4))
    if sem:
        r['semantic'] = torch.zeros(4, 19)
    if inten:
        r['intensity'] = torch.zeros(4)
    return [r]


def test_loss_index_follows_the_configuration():
    cfg = configs.nuscenes_single()
    _, _, index = train._supervision_setup(_rend(), cfg, 6000, 2)
    assert index == {'data': 0, 'depth': 1, 'sem': 2, 'int': 3, 'd_smo': 4, 's_smo': 5}
    _, _, index = train._supervision_setup(_rend(), cfg, 6000, 0)            # no patches: no smoothness terms
    assert index == {'data': 0, 'depth': 1, 'sem': 2, 'int': 3}
    _, _, index = train._supervision_setup(_rend(sem=False, inten=False), cfg, 6000, 2)
    assert index == {'data': 0, 'depth': 1, 'd_smo': 4}
    cfg.depth_loss = False
    rend, kc, index = train._supervision_setup(_rend(), cfg, 6000, 2)
    assert 'depth' not in index and kc['depth_mult'] == 0.
    assert rend['semantic'] is not None and rend['intensity'] is not None


def test_step_dependent_multipliers():
    """Z/train.py:330-371: depth 0.1 -> 0.4 and semantic 0.01 -> 0.04 after Config.end_step; both zero inside the
    pose-refinement window (start_step < step < 0.6 end_step)."""
    cfg = configs.nuscenes_single()
    assert cfg.pose_refine
    inside = (cfg.start_step + int(0.6 * cfg.end_step)) // 2
    for step, dep, sem in ((inside, 0., 0.), (int(0.6 * cfg.end_step) + 1, 0.1, 0.01), (cfg.end_step + 1, 0.4, 0.04)):
        _, kc, _ = train._supervision_setup(_rend(), cfg, step, 2)
        assert kc['depth_mult'] == dep and kc['sem_mult'] == sem, (step, kc)
        assert kc['int_mult'] == 0.1 and kc['smooth_mult'] == 0.01
    cfg.pose_refine = False
    _, kc, _ = train._supervision_setup(_rend(), cfg, inside, 2)
    assert kc['depth_mult'] == 0.1 and kc['sem_mult'] == 0.01


def test_fused_and_summed_paths_share_the_setup():
    """compute_losses and compute_losses_fused read the same (rendering, kernel configuration, index) triple."""
    import inspect
    for fn in (train.compute_losses, train.compute_losses_fused):
        assert '_supervision_setup(renderings, config, step, num_patch)' in inspect.getsource(fn)
