"""One reduced image (S PLMS steps, E alpha epochs) of bench.py's workload for `ncu` launch lists.

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/launches.csv python tools/ncu_step.py 5 1

The profiled region (cudaProfilerStart/Stop) is one generate() call AFTER a warm-up call captured the CUDA graphs, so
the list contains exactly the kernels of the timed region of bench.py, in the same proportions per UNet evaluation.
"""
import os
import sys
from pathlib import Path

os.environ.setdefault("STA_CUDNN_BENCHMARK", "0")  # no autotuner trial launches under the profiler

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200 import prompts as P  # noqa: E402
from diffusion_spacetime_attn_b200.pipeline import SpaceTimeAttnPipeline  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 5
E = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pipe = SpaceTimeAttnPipeline(steps=S, num_epochs=E, save_images=False)
items = [it for it in P.build_work_items(P.read_gpt(P.SYNTHETIC_GPT)) if len(it.object_names) == 2][:2]
conds = [pipe.to_device(pipe.encode([it])) for it in items]
pipe.generate([items[0]], conds[0])
torch.cuda.synchronize()
torch.cuda.profiler.start()
pipe.generate([items[1]], conds[1])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
