import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/input-inference-for-control_b200')
import numpy as np, i2c_b200
from i2c_b200 import capi
B,T=4096,200
rng=np.random.default_rng(0)
x0=np.array([np.pi,0.])+np.array([.3,.5])*rng.normal(size=(B,2)); mu_u=1e-2*rng.normal(size=(B,T,1))
Q,R=np.diag([1.,100.,1.]),np.diag([2.])
g=i2c_b200.BatchedI2c("PendulumKnown",B,T,Q,R,Q,100.,0.,mu_u,2*np.eye(1),x0=x0,max_iters=32)
g.run(5,capi.PH_LEARN,collect=False)
for name,ph in [("learn",capi.PH_LEARN),("fwd",capi.PH_FORWARD),("bwd",capi.PH_BACKWARD),("fwd+bwd",capi.PH_FORWARD|capi.PH_BACKWARD),("propagate",capi.PH_PROPAGATE)]:
    g.run(10,ph,collect=False); g.synchronize(); print(name, g.last_run_ms()/10)
