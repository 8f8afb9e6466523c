import sys, numpy as np
sys.path.insert(0,'/root/repo')
from openmoc_b200 import capi
from oracle.oracle_py import lib
x = np.concatenate([np.linspace(0, 30, 30001), np.logspace(-12, 3, 3001)])
ref = np.array([lib().moc_oracle_expF1(v) for v in x])
got = capi.eval_expF1(x)
print("max rel err double path vs oracle:", np.max(np.abs(got-ref)/ref))
