import sys, importlib
sys.path.insert(0, '/root/repo')
import zkb_loader
z = zkb_loader.load()
c = importlib.import_module("zkir_b200.circuits")
for window in (4096, 0):
    circ = c.random_circuit(1 << 20, 1024, c.GOLDILOCKS, 0x5EED0002, window=window)
    b = z.GpuBackend(0); b.set_field(c.GOLDILOCKS); b.push_gates(circ.gates, circ.const_pool); b.finalize(False)
    w = c.make_witnesses(circ, 1, seed=3)
    b.upload_inputs(None, w, 1)
    for i in range(3):
        print("run", i, "window", window, flush=True)
        b.run()
    print(b.timing())
