"""One grouped batch (4 x 1 M x 252, reference semantics) through optmc_price_american_batch: the ncu target for the
path kernel and the persistent sweep (development tool).  argv[1] = batch size, argv[2] = repetitions."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import options_model_b200  # noqa: E402,F401
from options_model_b200 import engine as E  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = E.Engine(0)
model = E.heston(100.0, 0.05, 1.0, v0=0.04, kappa=2.0, theta=0.04, xi=0.5, rho=-0.7)
for i in range(reps):
    p, se = eng.price_american_batch(model, 1_000_000, 100.0, 100.0, 1.0, np.full(B, 252), 1, "f32", E.RngSpec(seed=3 + i))
print(p, eng.kernel_times())
