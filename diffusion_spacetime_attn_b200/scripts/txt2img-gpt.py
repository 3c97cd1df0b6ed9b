#!/usr/bin/env python
"""Entrypoint with the reference's name and flags (scripts/txt2img-gpt.py); see _txt2img_common.py."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent))
from _txt2img_common import run  # noqa: E402

if __name__ == "__main__":
    sys.exit(run("gpt", "../../datasets/gpt.txt"))
