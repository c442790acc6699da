import sys, faulthandler
faulthandler.enable()
sys.path.insert(0, ".")
import cpnest_b200 as cpnest
import cpnest_b200.models as M
import tempfile
kind = sys.argv[1] if len(sys.argv) > 1 else "slice"
kw = dict(nslice=1) if kind == "slice" else (dict(nhamiltonian=1) if kind == "hmc" else {})
work = cpnest.CPNest(M.GaussianModel(2), verbose=1, nthreads=1, nlive=1000, maxmcmc=5000, poolsize=1000, output=tempfile.mkdtemp(), **kw)
print("constructed", flush=True)
work.run()
print("logZ", work.NS.logZ)
