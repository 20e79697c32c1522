"""Developer sweep of the tcgen05 stack's blocks-per-pass (BFCNN_KB_UMMA) on 4K frames; device-resident, CUDA events."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blind_image_denoising_b200 as bf

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 2
x = torch.randint(0, 256, (frames, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
out = torch.empty_like(x)
for kb in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["2", "1", "3"]):
    os.environ["BFCNN_KB_UMMA"] = kb
    m = bf.synthetic_model(18, precision="f16", pad_pow2=False)
    for _ in range(2):
        m(x, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m(x, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"kb={kb}: {ms:.3f} ms per {frames} frames -> {frames*2160*3840/1e3/ms:.0f} MP/s (stack {m.last_stack_ms():.2f} ms)", flush=True)
    m.close()
