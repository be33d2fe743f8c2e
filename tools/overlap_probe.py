"""Does splitting the batch over G handles (G streams, G host threads) on ONE GPU overlap the latency-bound
stages (Cholesky pivot chain) of one group with the bandwidth-bound ones of another?  Wall-clock probe."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def main():
    import torch
    import uvs_b200 as uvs
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    ws = bench.load_workload(B)
    opts = uvs.default_options(max_num_iterations=bench.K_LM, fixed_iterations=1)
    for G in (1, 2, 4):
        hs = []
        for g in range(G):
            s = uvs.Solver(0)
            s.upload(ws[g * B // G:(g + 1) * B // G], opts)
            hs.append(s)

        def run(s, n):
            for _ in range(n):
                s.reset_state(); s.solve()
        for s in hs:
            run(s, 3)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = [threading.Thread(target=run, args=(s, 10)) for s in hs]
        for t in th: t.start()
        for t in th: t.join()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        print("G=%d: %.2f ms/step  %.0f it/s" % (G, dt * 1e3, B * 10 / dt), flush=True)
        for s in hs:
            s.close()


main()
