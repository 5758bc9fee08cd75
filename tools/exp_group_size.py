"""Launch-group size experiment: ViT-256 over 16 regions with HB_VIT256_MAX_PATCHES sequences per launch sequence (the working
set per launch — xb + qkv + att = 1.97 MB per sequence — against the 126 MB L2).  One subprocess per size (the constant is read
at import); prints regions/s, mean SM clock and power.  python tools/exp_group_size.py [sizes...]"""
import json, os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import torch, pynvml
    from hipt_abmil_atec23_b200.hipt_4k import HIPT_4K
    from tests.common import seeded_modules
    pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
    dev = torch.device("cuda:0")
    m256, m4k = seeded_modules(0)
    hipt = HIPT_4K.from_modules(m256, m4k, dev, dev)
    R = int(sys.argv[2])
    regs = torch.randint(0, 256, (R, 3, 4096, 4096), dtype=torch.uint8, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
    for _ in range(3): hipt.forward_regions_u8(regs)
    torch.cuda.synchronize()
    samples, stop = [], threading.Event()
    def samp():
        while not stop.is_set():
            samples.append((pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            time.sleep(0.05)
    th = threading.Thread(target=samp); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 40
    e0.record()
    for _ in range(steps): out = hipt.forward_regions_u8(regs)
    e1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    s = samples[len(samples) // 3:] or samples
    print(json.dumps({"patches_per_launch": __import__("hipt_abmil_atec23_b200.engine", fromlist=["x"]).VIT256_MAX_PATCHES, "regions_per_s": R * steps / (e0.elapsed_time(e1) / 1e3),
                      "sm_mhz": sum(x[1] for x in s) / len(s), "power_w": sum(x[0] for x in s) / len(s), "checksum": float(out.double().sum())}))
else:
    sizes = [int(x) for x in sys.argv[1:]] or [512, 294, 221, 147, 73]
    for mp in sizes:
        env = dict(os.environ, HB_VIT256_MAX_PATCHES=str(mp))
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "16"], env=env, capture_output=True, text=True)
        print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:])
