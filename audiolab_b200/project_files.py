"""``ProjectFiles`` -- the boundary data carrier between pipeline stages.

Same shape as the reference's (/root/reference/util/data_classes.py:10-68): xxh64 of the input ->
``<output_path>/process/<name>_<hash8>/{source,stems,...}``; ``add_output(process, paths)`` records
``last_outputs``.
"""
from __future__ import annotations

import os
from shutil import copyfile
from typing import List, Union

import xxhash

output_path = os.environ.get("AUDIOLAB_OUTPUT_PATH", os.path.join(os.getcwd(), "outputs"))


class ProjectFiles:
    def __init__(self, input_file: str, root: str = None):
        h = xxhash.xxh64()
        with open(input_file, "rb") as f:
            while chunk := f.read(1 << 16):
                h.update(chunk)
        self.file_hash = h.hexdigest()[:8]
        name, _ = os.path.splitext(os.path.basename(input_file))
        self.project_dir = os.path.join(root or output_path, "process", f"{name}_{self.file_hash}")
        source_dir = os.path.join(self.project_dir, "source")
        os.makedirs(source_dir, exist_ok=True)
        self.src_file = os.path.join(source_dir, os.path.basename(input_file))
        if not os.path.exists(self.src_file):
            copyfile(input_file, self.src_file)
        self.last_outputs: List[str] = []
        self.video_sources = {}
        self.file_dict = {"source": [self.src_file]}
        self.output_dict = {}
        for root_, _dirs, files in os.walk(self.project_dir):
            if root_ == self.project_dir:
                continue
            folder = os.path.basename(root_)
            bucket = self.file_dict.setdefault(folder, [])
            for fn in files:
                p = os.path.join(root_, fn)
                if p not in bucket:
                    bucket.append(p)

    def add_output(self, process: str, outputs: Union[List[str], str]):
        if isinstance(outputs, str):
            outputs = [outputs]
        self.last_outputs = outputs
        self.file_dict.setdefault(process, []).extend(outputs)
        self.output_dict.setdefault(process, []).extend(outputs)

    def all_outputs(self) -> List[str]:
        seen: List[str] = []
        for key, files in self.output_dict.items():
            if key in ("merge", "convert", "export"):
                continue
            for f in files:
                if os.path.exists(f) and f not in seen:
                    seen.append(f)
        return seen
