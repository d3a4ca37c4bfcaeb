timeout 300 python -m pytest tests/test_gpu_unet.py -q 2>&1 | tail -4
timeout 200 python tools/unet_vs_torch.py 2>&1 | tail -3
