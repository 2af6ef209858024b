#!/usr/bin/env python
"""Convert the reference's pre-computed WebDataset embedding shards (.tar: per-sample json + pickled tensors,
thinkdiff/tasks/image_text_process_data.py:94-118) into flat .tdemb shards (thinkdiff_mlre_b200/shards.py):
    python scripts/convert_shards.py --out /data/ccsbu/00000 /data/ccsbu_wds/00000.tar [more.tar ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from thinkdiff_mlre_b200.shards import _main  # noqa: E402

if __name__ == "__main__":
    _main()
