"""Short driver for ncu captures of the HBM-side kernels at 262144 rays: sampler iteration, compositor fwd/bwd (canonical +
fast), MVS cost lookup (GPU box)."""
import os, sys, warnings
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
import runpy
sys.argv = ['bench_hbm_kernels.py', '262144']
import tools.bench_hbm_kernels  # noqa: F401  (runs the micro-benchmark once; ncu picks the launches it is told to)
