#!/usr/bin/env python
"""JPEG decode throughput: the bench.py `decode_jpeg` key on its own.  usage: jpeg_bench.py [n_files] [distinct]"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from bbocr_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
d = int(sys.argv[2]) if len(sys.argv) > 2 else 8
print(json.dumps(bench.decode_jpeg(_lib.Handle(0), bench.peaks()[2], n, d)))
