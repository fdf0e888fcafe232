"""Developer experiment: throughput of SA1-sized FPS launches issued round-robin on D streams."""
import importlib, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("3dvlp_b200")
syn = importlib.import_module("3dvlp_b200.synthetic")
x = syn.make_batch(0, 8, 40000, 0).cuda().contiguous()
for depth in (1, 2, 3, 4):
    streams = [torch.cuda.Stream() for _ in range(depth)]
    for _ in range(2):
        pkg._ext.furthest_point_sampling_gather(x, 2048)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record()
    for st in streams:
        st.wait_stream(main)
    n = 24
    for i in range(n):
        with torch.cuda.stream(streams[i % depth]):
            pkg._ext.furthest_point_sampling_gather(x, 2048)
    for st in streams:
        main.wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
    print("depth %d: %.3f ms per 8-scene FPS (cluster=%s)" % (depth, e0.elapsed_time(e1) / n, os.environ.get("PN2_FPS_CLUSTER", "auto")))
