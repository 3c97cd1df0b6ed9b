"""Experiment: two prompts in flight on one GPU (two pipeline objects, two host threads, two CUDA streams), each UNet evaluation
still batch 1 — the reference's own way to use a big GPU is several processes per card (process_id.py).  Prints images/s for
1 and 2 concurrent prompt streams."""
import sys
import threading
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200 import native, prompts as P  # noqa: E402
from diffusion_spacetime_attn_b200.pipeline import SpaceTimeAttnPipeline  # noqa: E402

N_IMG = int(sys.argv[1]) if len(sys.argv) > 1 else 2
items = P.build_work_items(P.read_gpt(P.SYNTHETIC_GPT))


def worker(pipe, stream, my_items, out):
    with torch.cuda.stream(stream):
        for it in my_items:
            pipe.generate([it], pipe.encode([it]), check_device_error=False)
        stream.synchronize()
    out.append(len(my_items))


def run(n_streams):
    import os

    os.environ["STA_MAX_SLOTS"] = "64"
    pipes = [SpaceTimeAttnPipeline(device="cuda", seed=0, steps=50, num_epochs=3, save_images=False) for _ in range(n_streams)]
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    # warm-up: one image per pipeline, ONE AFTER THE OTHER (CUDA-graph capture must not overlap another thread's CUDA calls)
    for k in range(n_streams):
        done = []
        worker(pipes[k], streams[k], [items[k * 8 + j] for j in range(N_IMG)], done)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        done, threads = [], []
        for k in range(n_streams):
            mine = [items[k * 8 + j] for j in range(N_IMG)]
            th = threading.Thread(target=worker, args=(pipes[k], streams[k], mine, done))
            th.start()
            threads.append(th)
        for th in threads:
            th.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f"streams={n_streams} rep={rep}: {sum(done)} images in {dt:.2f} s -> {sum(done) / dt:.3f} images/s; "
              f"reserved {torch.cuda.memory_reserved() / 2**30:.1f} GiB", flush=True)
    del pipes
    import gc

    gc.collect()
    torch.cuda.empty_cache()


if __name__ == "__main__":
    run(1)
    run(2)
    print("device_error", native.device_error())
