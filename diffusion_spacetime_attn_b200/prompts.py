"""Prompt-set readers for the three entrypoints and the layout sidecar format.

Formats (reference scripts/txt2img-{gpt,mscoco,vsr}.py:254-261 and datasets/*.txt):
  gpt     4-line records  `Objects: a, b` / `Relation: ...` / `Sentence: ...` / blank; the prompt is line 4i+2
          without its 10-character "Sentence: " prefix (txt2img-gpt.py:261)
  mscoco  one caption per line
  vsr     one sentence per line ("The bed is below the cat.")
Layouts: the reference calls the layout predictor per prompt (`inference_sentence`, txt2img-gpt.py:307), which needs
fairseq/spaCy and a checkpoint that do not exist offline.  A JSON sidecar {prompt: {phrase: [x, y]}} — the
predictor's own output format, inference_coco.py:535-544 — can be supplied with --bboxes-from; without it objects
come from the `Objects:` line (gpt) or a noun heuristic, and positions from pipeline.synthetic_layout.
"""
from __future__ import annotations

import json
import re
from pathlib import Path
from typing import Dict, List, Optional, Tuple

from .pipeline import WorkItem, synthetic_layout

DATA_DIR = Path(__file__).resolve().parent / "data"
SYNTHETIC_GPT = DATA_DIR / "synthetic_gpt.txt"


def read_gpt(path, limit: int = 500) -> List[Tuple[str, List[str]]]:
    rows = Path(path).read_text().split("\n")[: 4 * limit]
    out = []
    for i in range(min(limit, len(rows) // 4)):
        objects = [o.strip() for o in rows[4 * i][len("Objects:"):].split(",") if o.strip()]
        out.append((rows[4 * i + 2][10:], objects))
    return out


_STOP = {"the", "a", "an", "is", "are", "of", "to", "and", "with", "on", "in", "at", "this", "that", "there", "it",
         "left", "right", "above", "below", "front", "behind", "next", "near", "as", "has", "was", "by"}


def guess_objects(sentence: str, max_objects: int = 5) -> List[str]:
    """Very small stand-in for the reference's spaCy noun chunks ∩ COCO categories (inference_coco.py:518-528)."""
    words = [w for w in re.findall(r"[a-z]+", sentence.lower()) if w not in _STOP and len(w) > 2]
    seen, out = set(), []
    for w in words:
        if w not in seen:
            seen.add(w)
            out.append(w)
    return out[:max_objects] or ["object"]


def read_lines(path, limit: int = 500) -> List[Tuple[str, List[str]]]:
    lines = [l.strip() for l in Path(path).read_text().split("\n") if l.strip()][:limit]
    return [(l, guess_objects(l, 2 if "vsr" in str(path) else 5)) for l in lines]


def load_layouts(path: Optional[str]) -> Dict[str, Dict[str, List[float]]]:
    return json.loads(Path(path).read_text()) if path else {}


def build_work_items(records, layouts=None, start: int = 0, seed: int = 1, min_objects: int = 0,
                     max_objects: int = 8, force_objects: Optional[int] = None) -> List[WorkItem]:
    layouts = layouts or {}
    items = []
    for idx, (prompt, objects) in enumerate(records):
        if prompt in layouts:  # the predictor's output: {phrase: [x, y]}
            names = list(layouts[prompt].keys())
            boxes = [list(map(float, layouts[prompt][k])) for k in names]
        else:
            names = list(objects)[:max_objects]
            if force_objects is not None:
                names = (names + [f"object {k}" for k in range(force_objects)])[:force_objects]
            boxes = synthetic_layout(names, prompt)
        if len(names) < min_objects:
            continue
        items.append(WorkItem(prompt=prompt, prompt_idx=start + idx, object_names=names, bboxes=boxes, seed=seed))
    return items
