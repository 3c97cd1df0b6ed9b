"""DDIM sampler WITH the attention injection and alpha optimisation — an EXTENSION, not reference behaviour.

The reference's ddim.py is vanilla upstream code that calls `apply_model` with one positional argument short
(ddim.py:172,177 -> ddpm.py:1005 -> ddpm.py:1420: TypeError), i.e. DDIM has no working injection path in the fork
(SURVEY.md §0).  BASELINE.json configs[4] asks for "100 DDIM steps"; it is defined here by analogy with
p_sample_plms: the same CFG batch through `apply_model_extra` with coef = weighting_parameter[:, i], followed by the
eta = 0 DDIM update x_prev = sqrt(a_prev) * pred_x0 + sqrt(1 - a_prev) * e_t (one UNet evaluation per step).
"""
from .plms import PLMSSampler


class DDIMSampler(PLMSSampler):
    method = "ddim"
