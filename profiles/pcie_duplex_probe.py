"""PCIe duplex probe: 2 MiB uploads and 2.28 MiB downloads (one stereo stream of the bench workload) issued
on two streams at once, pinned memory -- the floor of the host-pointer path per stream.  profiles/ aid."""
import time
import torch

d = torch.device("cuda")
nin, nout = 262144 * 2, 285350 * 2
K = 16
hin = [torch.empty(nin, dtype=torch.float32).pin_memory() for _ in range(K)]
hout = [torch.empty(nout, dtype=torch.float32).pin_memory() for _ in range(K)]
gin = [torch.empty(nin, dtype=torch.float32, device=d) for _ in range(K)]
gout = [torch.empty(nout, dtype=torch.float32, device=d) for _ in range(K)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, reps=20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(K):
            if up:
                with torch.cuda.stream(s1):
                    gin[i].copy_(hin[i], non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    hout[i].copy_(gout[i], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps / K * 1e6


for _ in range(2):
    run(True, True, 2)
u, dn, both = run(True, False), run(False, True), run(True, True)
print(f"per stream: upload alone {u:.1f} us ({nin * 4 / u / 1e3:.1f} GB/s), download alone {dn:.1f} us ({nout * 4 / dn / 1e3:.1f} GB/s), "
      f"both at once {both:.1f} us -> floor {nout / both / 1e3:.2f} Gsamples/s")
