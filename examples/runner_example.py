# torchrun --standalone --local-addr 127.0.0.1 --nproc-per-node 4 examples/runner_example.py
# (the reference: mpirun -n 4 python -m runner_example)
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from pyDNMFk.runner import pyDNMFk_Runner  # noqa: E402

runner = pyDNMFk_Runner(itr=100, init='nnsvd', verbose=True, norm='fro', method='mu', precision=np.float32,
                        checkpoint=False, sill_thr=0.6)
golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden') + '/'
results = runner.run(grid=[4, 1], fpath=golden, fname='wtsi_X', ftype='npy', results_path='results/', k_range=[1, 3], step_k=1)
W = results["W"]
H = results["H"]
