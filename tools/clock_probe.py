"""SM clock / power under a sustained load of each F16 engine (3 s each): python tools/clock_probe.py"""
import sys, os, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import blind_image_denoising_b200 as bf

def sample(stop, out):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu,clocks_event_reasons.active",
                          "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=subprocess.PIPE, text=True)
    for ln in p.stdout:
        out.append(ln.strip())
        if stop.is_set():
            break
    p.terminate()

x = torch.randint(0, 256, (4, 2160, 3840, 3), dtype=torch.uint8, device="cuda")
o = torch.empty_like(x)
m = bf.synthetic_model(18, precision="f16", pad_pow2=False)
for env in ("0", "1", "0"):
    os.environ["BFCNN_UMMA_REGIONS"] = env
    stop, lines = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, lines)); th.start()
    time.sleep(0.3)
    m(x, out=o); torch.cuda.synchronize()
    t0 = time.time(); n = 0
    while time.time() - t0 < 3.0:
        m(x, out=o); n += 1
        if n % 4 == 0: torch.cuda.synchronize()
    torch.cuda.synchronize()
    dt = time.time() - t0
    stop.set(); th.join()
    print(f"REGIONS={env}: {4*2160*3840*n/1e6/dt:.0f} MP/s over {dt:.1f}s; samples (sm, max, W, limit, C, reasons):")
    for ln in lines[5:-2:4]:
        print("   ", ln)
