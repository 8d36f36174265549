"""per-layer eager stage times of the engine (CUDA events): python tools/layer_times.py [batch]"""
import os
import sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "second.pytorch_b200"))
import torch
import bench
from b2second import config, loader, models
from b2second.engine import InferenceEngine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sp = loader.product_spconv()
cfg = config.get_config("car.fhd")
net = models.build_network(cfg, sp).eval()
models.synthetic_weights_(net, "car.fhd", seed=0)
net = net.cuda()
eng = InferenceEngine(net, batch_size=B, max_points=30000, use_cuda_graph=False)
clouds = [torch.from_numpy(c).cuda() for c in bench.make_clouds("car.fhd", B, 29000)]
eng.load_points(clouds)
st = eng.run_timed(iters=5)
stats = {s["index"]: s for s in eng.sparse_layer_stats()}
for k, v in st.items():
    extra = ""
    if k.startswith("sparse_conv"):
        s = stats[int(k[len("sparse_conv"):])]
        extra = "  cin %d cout %d K %d rows %d pairs %d tc %s" % (s["cin"], s["cout"], s["K"], s["n_out"], s["pairs"],
                                                                  eng.layers[s["index"]]["tc"])
    print("%-16s %8.1f us%s" % (k, v * 1e3, extra))
