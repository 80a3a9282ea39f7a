import os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def load_golden(name):
    with np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz')) as z:
        return {k: z[k] for k in z.files}
