"""Throughput of the preprocessing kernels vs the reference's CPU loader arithmetic (cv2 + numpy on one host core)."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import ammcnet_aaai2021_b200 as A
dev = "cuda:0"
peaks = {}
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm = float(peaks.get("hbm_gbs", 6650.0))
for name, (h0, w0) in (("ped2", (240, 360)), ("avenue", (360, 640)), ("shanghaitech", (480, 856))):
    n = 256
    bgr = torch.randint(0, 256, (n, h0, w0, 3), dtype=torch.uint8, device=dev)
    flow = torch.randn((n, h0, w0, 2), device=dev)
    for fn, src, out_c, label in ((A.preprocess_frames, bgr, 3, "frames"), (A.preprocess_flow, flow, 2, "flow")):
        for _ in range(3): fn(src)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn(src)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        # algorithmic bytes: every source byte once + every output byte once (flow: channel 0 of the source only)
        src_bytes = src.numel() * src.element_size() // (2 if label == "flow" else 1)
        bytes_ = src_bytes + n * out_c * 256 * 256 * 4
        print(json.dumps({"dataset": name, "what": label, "frames": n, "ms": ms, "frames_per_s": n / ms * 1e3,
                          "algorithmic_GBps": bytes_ / ms / 1e6, "frac_of_hbm_peak": bytes_ / ms / 1e6 / hbm}), flush=True)
# CPU reference arithmetic for one frame (cv2 when available, else the oracle's numpy restatement)
try:
    import cv2
    img = np.random.default_rng(0).integers(0, 256, (360, 640, 3), dtype=np.uint8)
    t0 = time.perf_counter()
    for _ in range(200):
        r = cv2.resize(cv2.cvtColor(img, cv2.COLOR_BGR2RGB), (256, 256)).astype(np.float32) / 255.0
        r = (r.transpose(2, 0, 1) - 0.5) / 0.5
    dt = (time.perf_counter() - t0) / 200
    print(json.dumps({"cpu_one_core_frames_per_s_avenue": 1.0 / dt, "kind": "cv2 + numpy, one host core"}))
except Exception as e:
    print(json.dumps({"cpu": "cv2 unavailable: %s" % e}))
