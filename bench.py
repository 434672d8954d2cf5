#!/usr/bin/env python
"""Benchmark of the SlowTV-monodepth training hot path on B200 (one JSON line on stdout from rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c5] [--api auto|plugin|native]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one full training step of a BASELINE.json configuration — default configs[2] (ConvNeXt-T depth + ResNet-18 pose,
384x640, batch 8 per GPU, 2 support frames, 4 scales, min-reprojection + automask + edge-aware smoothness, AdamW), the one the
metric is quoted on; --config selects the others: networks forward, fused loss, backward, gradient all-reduce (N > 1),
optimiser step. Synthetic video triplets, random-init weights. By default the step is driven THROUGH THE REFERENCE'S OWN MODULE
(`MonoDepthModule(cfg).step`, built by its registry after `plugin.install()`) when the reference build (oracle/_ref) is present.

  value  images/s over all ranks, inputs resident in HBM before the timed region (CUDA events, max over ranks)
  e2e    the same metric through the public API with HOST (pinned) batches: H2D copy of every step's batch and a D2H
         read of the loss inside the timed region
  roofline      fused photometric loss (forward + backward entry points), algorithmic bytes (148 + 164 B per
                target pixel for n=2, S=4) over the live CUDA-event duration of those calls, vs MEASURED_PEAKS.json
  roofline_tensor  all tcgen05 products of the step vs the cuBLAS TF32 peak measured on this GPU in the same run
  gpu_torch_baseline  the library path (torch.nn + ATen: cuDNN / cuBLAS TF32) of the same step on the same GPU
  cpu_baseline  the reference's own step on the host cores (bounded sample)

`--impl reference` times the reference's own CPU step alone: its unmodified modules from the byte-compiled build oracle/_ref
(the source tree does not exist on the GPU box), or the oracle port when that build is absent; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_SCALES = 4
# BASELINE.json configs (SURVEY 8d "Configs -> concrete shapes"); gflop = forward GFLOP per image (x3 for forward + both gradients).
CONFIGS = {
    'c2': dict(depth='resnet18', pose='resnet18', shape=(192, 640), batch=12, n=2, learn_K=False, gflop=8.9 + 7.1 + 2*10.3,
               label='configs[1]: ResNet-18 depth + ResNet-18 pose, 192x640, batch 12, 2 support frames, S=4'),
    'c3': dict(depth='convnext_tiny', pose='resnet18', shape=(384, 640), batch=8, n=2, learn_K=False, gflop=43.6 + 14.0 + 2*20.7,
               label='configs[2]: ConvNeXt-T depth + ResNet-18 pose (KBR default), 384x640, batch 8/GPU, 2 support frames, S=4'),
    'c4': dict(depth='convnext_tiny', pose='resnet18', shape=(384, 640), batch=8, n=4, learn_K=True, gflop=43.6 + 14.0 + 4*20.7,
               label='configs[3]: learned intrinsics + 4-scale loss + auto-mask, ConvNeXt-T, 384x640, batch 8/GPU, 4 support frames'),
    'c5': dict(depth='convnext_base', pose='resnet18', shape=(512, 1024), batch=4, n=2, learn_K=False, gflop=320.9 + 34.1 + 2*44.1,
               label='configs[4]: HR stress, ConvNeXt-B depth, 512x1024, batch 4/GPU, 2 support frames, S=4'),
}
CPU_BATCH = 2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c3', choices=sorted(CONFIGS), help='BASELINE.json configuration (default c3 = configs[2], the one the metric is quoted on)')
    ap.add_argument('--batch', type=int, default=0, help='per-GPU batch (default: the configuration\'s)')
    ap.add_argument('--api', default='auto', choices=['auto', 'plugin', 'native'],
                    help="plugin: the reference's own MonoDepthModule built through its registry after plugin.install() (needs oracle/_ref); "
                         'native: slowtv_monodepth_b200.trainer.MonoDepthStep; auto: plugin when the reference build is present')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-torch-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='enqueue every step eagerly instead of replaying the captured CUDA graph')
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    args.cfg = cfg
    if args.batch <= 0: args.batch = cfg['batch']
    return args


def peaks() -> tuple[float, str]:
    f = ROOT/'MEASURED_PEAKS.json'
    if f.is_file():
        return float(json.loads(f.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference step
# ---------------------------------------------------------------------------------------------------------------------
def reference_cfg(batch: int, c: dict) -> dict:
    """The reference's own experiment configuration of the benchmarked workload: cfg/abl_learn_K/default.yaml (ConvNeXt-T depth +
    ResNet-18 pose) with the KBR loss / depth-range / optimiser settings (cfg/kbr/default.yaml), random init."""
    from oracle import ref_shim
    cfg = ref_shim.load_cfg('abl_learn_K/default.yaml')
    kbr = ref_shim.load_cfg('kbr/default.yaml')
    cfg['net']['depth'].update(enc_name=c['depth'], pretrained=False)
    cfg['net']['pose'].update(enc_name=c['pose'], pretrained=False, learn_K=c['learn_K'])
    cfg['loss'] = {k: kbr['loss'][k] for k in ('img_recon', 'disp_smooth')}
    cfg['optimizer'] = {'type': 'adamw', 'lr': 1e-4, 'weight_decay': 1e-3}
    cfg['scheduler'] = None
    cfg['loader'] = {'batch_size': batch}
    cfg['dataset'] = {}
    cfg['trainer'].update(min_depth=0.1, max_depth=100, always_fwd_pose=False, aspect_ratio_aug_prob=0.0)
    return cfg


def cpu_reference(c: dict, steps: int, warmup: int, batch: int = CPU_BATCH, budget_s: float = 200.0) -> dict:
    """The reference step on the host cores. With the reference build present (oracle/_ref, or /root/reference in the build
    container) this is the REFERENCE'S OWN code — `MonoDepthModule(cfg).step(batch)` built through its registry / parsers, its
    own handlers / ViewSynth / losses, `loss.backward()`, the AdamW its `parsers.get_opt` builds — with only the absent third-party
    packages stubbed (timm encoders restated in oracle/nets.py, Lightning as a plain nn.Module). Otherwise the oracle port.
    BASELINE.md section 3 protocol: warm-up steps, then the MEDIAN of the timed steps; per-image rate at a reduced batch."""
    import statistics
    import warnings
    from oracle import ref_shim
    from slowtv_monodepth_b200 import synthetic as syn
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    batches = [syn.make_batch(batch, c['n'], c['shape'], seed=s) for s in range(2)]
    if ref_shim.available():
        warnings.filterwarnings('ignore')
        ref_shim.load()
        import src.core.trainer as rt
        from src.tools import parsers
        cfg = reference_cfg(batch, c)
        module = rt.MonoDepthModule(cfg).train()
        if not torch.cuda.is_available():  # build container only: the module's timers call torch.cuda.synchronize()
            from src.utils import MultiLevelTimer
            module.timer = MultiLevelTimer(name='MonoDepthModule', as_ms=True, precision=4, sync_gpu=False)
        opt = parsers.get_opt(module.nets, dict(cfg['optimizer']))

        def train_step(b):
            opt.zero_grad(set_to_none=True)
            loss = module.step(b, mode='train')[0]
            loss.backward()
            opt.step()
        kind = 'reference'
        what = f"the reference's own MonoDepthModule.step + backward + AdamW ({ref_shim.kind()} build; timm/Lightning stubbed)"
    else:
        from oracle.step import OracleTrainer
        tr = OracleTrainer(c['depth'], c['pose'], learn_K=c['learn_K']).train()
        train_step = tr.train_step
        kind = 'port'
        what = 'oracle port of the reference step (oracle/step.py)'
    t_start = time.perf_counter()
    for i in range(warmup): train_step(batches[i % 2])
    times = []
    for i in range(steps):
        t0 = time.perf_counter()
        train_step(batches[i % 2])
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and len(times) >= 3: break   # bounded sample: never run past the budget
    med = statistics.median(times)
    return {'value': batch/med, 'ms_per_step': med*1e3, 'cores': cores, 'batch': batch, 'kind': kind, 'steps_done': len(times),
            'sample': f'median of {len(times)} timed step(s) after {warmup} warm-up at batch {batch} of the same workload: {what}, '
                      f'PyTorch {torch.__version__} CPU fp32, {cores} threads'}


def run_reference_arm(args, out) -> None:
    if int(os.environ.get('RANK', '0')) != 0: return
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    r = cpu_reference(args.cfg, steps, warmup)
    line = {
        'impl': 'reference', 'metric': 'training images/sec', 'value': round(r['value'], 4), 'unit': 'images/s',
        'n_gpus': args.gpus, 'steps': r['steps_done'], 'warmup': warmup, 'ms_per_step': round(r['ms_per_step'], 2),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': args.cfg['label'], 'cpu_batch_per_step': r['batch'], 'requested_steps': args.steps,
                   'requested_warmup': args.warmup,
                   'note': 'per-image rate at a reduced batch (BASELINE.md section 3); median of the timed steps; the run stops early only past a 200 s budget'},
        'cpu_baseline': {'value': round(r['value'], 4), 'unit': 'images/s', 'cores': r['cores'], 'kind': r['kind'], 'sample': r['sample']},
        'e2e': {'value': round(r['value'], 4), 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), file=out, flush=True)


# ---------------------------------------------------------------------------------------------------------------------
# Clock sampling during the timed region
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout: self.rows.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None: return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try: self.proc.wait(timeout=2)
        except Exception: self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [c.strip() for c in r.split(',')]
            if len(f) < 7: continue
            try: sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError: continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith('active'): reasons.add(nm)
        if not sm: return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['no samples']}
        sm.sort()
        return {'sm_mhz': sm[len(sm)//2], 'sm_max_mhz': max(mx), 'power_w_max': max(pw), 'samples': len(sm), 'reasons': sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------------
# Our arm
# ---------------------------------------------------------------------------------------------------------------------
def _claim_stdout():
    """The contract is ONE JSON line on stdout. Libraries write banners there from C (e.g. "NCCL version ..." on the first
    collective), so file descriptor 1 is pointed at stderr for the whole run and the JSON line goes to the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, 'w')


def measured_tf32_peak(dev) -> float:
    """TFLOP/s of cuBLAS TF32 on this GPU, measured now: torch.matmul fp32 8192^3 with TF32 enabled, best of 6 (CUDA events)."""
    n = 8192
    # (no torch RNG here: the default CUDA generator is registered with the captured step graph)
    a = (torch.arange(n*n, device=dev, dtype=torch.float32).reshape(n, n) % 251.0)/251.0 - 0.5
    b = (torch.arange(n*n, device=dev, dtype=torch.float32).reshape(n, n) % 241.0)/241.0 - 0.5
    best = float('inf')
    for i in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b); e1.record()
        torch.cuda.synchronize()
        if i >= 2: best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2.0*n**3/1e12/(best/1e3)


def photo_traffic() -> tuple[float | None, str]:
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/r2_photo_traffic.json, written by
    tools/ncu_traffic.py from `ncu --set full`), or None when no capture of this build has been committed."""
    f = ROOT/'profiles'/'r2_photo_traffic.json'
    if not f.is_file(): return None, 'no committed ncu capture (profiles/r2_photo_traffic.json)'
    d = json.loads(f.read_text())
    return float(d['dram_bytes_per_step']), d.get('source', str(f.name))


def torch_gpu_baseline(c: dict, batch: int, dev, steps: int = 10, warmup: int = 3) -> dict:
    """The library path on the same box: the oracle port of the reference step (plain torch.nn modules + ATen ops: cuDNN / cuBLAS
    TF32, cudnn.benchmark, torch.optim.AdamW) run eagerly on this GPU — what the reference itself would execute here."""
    from oracle.step import OracleTrainer
    from slowtv_monodepth_b200 import synthetic as syn
    torch.manual_seed(0)
    tr = OracleTrainer(c['depth'], c['pose'], learn_K=c['learn_K']).to(dev).train()
    bs = [syn.make_batch(batch, c['n'], c['shape'], seed=900 + s, device=dev) for s in range(2)]
    for i in range(warmup): tr.train_step(bs[i % 2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps): tr.train_step(bs[i % 2])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/steps
    del tr, bs
    torch.cuda.empty_cache()
    return {'value': round(batch/(ms/1e3), 2), 'unit': 'images/s', 'ms_per_step': round(ms, 3), 'steps': steps,
            'what': 'oracle port of the reference step as torch.nn modules + ATen ops on this GPU (cuDNN/cuBLAS TF32, cudnn.benchmark, '
                    'torch.optim.AdamW), eager, inputs resident'}


def main() -> None:
    args = parse()
    out = _claim_stdout()
    if args.impl == 'reference': return run_reference_arm(args, out)

    import torch.distributed as dist
    from oracle import ref_shim
    from slowtv_monodepth_b200 import _lib as L, functional as F_, synthetic as syn
    from slowtv_monodepth_b200.optim import FlatAdamW
    from slowtv_monodepth_b200.trainer import GraphedTrainStep, MonoDepthStep, default_cfg, gradient_buckets

    c = args.cfg
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank, local = int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available(): raise SystemExit('bench.py needs a CUDA device (no CPU fallback for --impl ours).')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    # Reference numerics: fp32 storage, TF32 tensor-core matmuls, cudnn autotune (cfg/default.yaml:170-174, trainer.py:30).
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    torch.set_float32_matmul_precision('high')

    L.lib()  # fail loudly if libstv.so is missing
    b, (H, W), n_supp = args.batch, c['shape'], c['n']
    torch.manual_seed(1234)  # identical initial weights on every rank (FlatAdamW also broadcasts rank 0's, like DDP)
    use_plugin = args.api == 'plugin' or (args.api == 'auto' and ref_shim.available())
    if use_plugin:
        # The drop-in: the REFERENCE'S OWN module, built by its own parsers through its registry after plugin.install(); its own
        # step / forward_loss / handlers call sites run, with the B200 classes behind them.
        import warnings
        warnings.filterwarnings('ignore')
        from slowtv_monodepth_b200 import plugin
        ref_shim.load()
        plugin.install()
        import src.core.trainer as rt
        model = rt.MonoDepthModule(reference_cfg(b, c)).to(dev).train()
        api = f"plugin: the reference's MonoDepthModule(cfg).step via src.registry after plugin.install() ({ref_shim.kind()} build of the reference)"
    else:
        model = MonoDepthStep(default_cfg(c['depth'], c['pose'], learn_K=c['learn_K'])).to(dev).train()
        api = 'native: slowtv_monodepth_b200.trainer.MonoDepthStep.step'
    opt = FlatAdamW(model.nets, lr=1e-4, weight_decay=1e-3, buckets=gradient_buckets(model.nets))
    n_params = opt.flat.numel()

    # Distinct batches per rank (DistributedSampler equivalent): seed = base + rank. Two batches are rotated.
    host = [syn.make_batch(b, n_supp, c['shape'], seed=100*rank + s, pin=True) for s in range(2)]
    to_dev = lambda bt: ({k: (v.to(dev, non_blocking=True) if k != 'supp_idxs' else v) for k, v in bt[0].items()},
                         {k: v.to(dev, non_blocking=True) for k, v in bt[1].items()}, {})
    resident = [to_dev(bt) for bt in host]
    h2d_bytes = sum(v.numel()*v.element_size() for d in host[0][:2] for k, v in d.items() if k != 'supp_idxs')

    def eager_step(batch):
        opt.zero_grad()
        loss = model.step(batch)[0]
        loss.backward()
        opt.all_reduce_async()
        opt.step()
        return loss

    for i in range(3): eager_step(resident[i % 2])   # first-use initialisation; also the photometric kernels' live timing below
    # Per-entry-point device times (roofline figures): CUDA events around each libstv call on the launching stream. The step
    # normally runs the pose network and the weight gradients on auxiliary streams beside the main chain; for these per-kernel
    # figures everything is enqueued on ONE stream, so that an event pair brackets exactly its own kernels.
    aux_cfg = (F_.WGRAD_STREAM, F_.BRANCH_STREAMS)
    F_.WGRAD_STREAM = F_.BRANCH_STREAMS = False
    F_.enable_kernel_timing(True)
    eager_step(resident[1])
    torch.cuda.synchronize()
    F_.reset_kernel_timings()
    TIMED_STEPS = 5
    # The eager loop is host-bound (~30 ms of enqueue work for ~18 ms of kernels): without a head start the device waits for the
    # host INSIDE an entry point's event pair (several launches per call) and the figure depends on the box's CPU. A device-side
    # spin in front of every timed step lets the host enqueue the whole step first; the kernels then run back to back.
    spin_cycles = int(0.045*1.9e9)
    for i in range(TIMED_STEPS):
        torch.cuda._sleep(spin_cycles)
        eager_step(resident[i % 2])
    torch.cuda.synchronize()
    kt = F_.kernel_timings()
    F_.enable_kernel_timing(False)
    F_.WGRAD_STREAM, F_.BRANCH_STREAMS = aux_cfg
    for i in range(2): eager_step(resident[i % 2])
    graphed, graph_note = None, 'eager (--no-graph)'
    if not args.no_graph:
        try:
            graphed = GraphedTrainStep(model, opt, resident[0])
            graph_note = 'CUDA graph replay of fwd+loss+bwd' + (' (pose network and weight gradients on auxiliary streams inside the graph)' if F_.BRANCH_STREAMS or F_.WGRAD_STREAM else '') + (' + per-bucket NCCL all-reduce overlapped inside the graph' if graphed.overlap
                                                                  else (', then all-reduce' if world > 1 else '')) + ', then AdamW'
            if world > 1 and not graphed.overlap and hasattr(graphed, 'overlap_error'): graph_note += f' (overlap capture failed: {graphed.overlap_error})'
        except Exception as e:  # keep the benchmark alive: report the eager number and say why
            torch.cuda.synchronize()
            graph_note = f'eager (graph capture failed: {type(e).__name__}: {str(e)[:120]})'
    train_step = eager_step if graphed is None else graphed.run

    def sync_all():
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; CUDA events on the launching stream; max over ranks (ms)."""
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for i in range(steps): fn(i)
        timed.host_ms = (time.perf_counter() - t0)*1e3/steps  # host time to ENQUEUE one step (launch-bound when close to ms_per_step)
        e1.record()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    for i in range(max(args.warmup, 3)): train_step(resident[i % 2])

    # ---- value: device-resident inputs ------------------------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0: sampler.start()
    launches0 = L.launch_count()
    eager_step(resident[0])
    launches_per_step = L.launch_count() - launches0  # libstv kernels in one step (a graph replay launches the same kernels)
    ms = timed(lambda i: train_step(resident[i % 2]), args.steps)
    host_ms = timed.host_ms
    launches = launches_per_step*args.steps
    clocks = sampler.stop() if rank == 0 else {}

    # ---- e2e: host batches through the public API -------------------------------------------------------------------
    last = {}

    def e2e_step(i):
        # Every step: H2D of a batch from pinned host memory into the step's input buffers, and a D2H read of the step's loss.
        # With the captured graph the input path is pipelined like a data loader's: batch i+1 crosses PCIe on a copy stream
        # while step i computes (both copies are inside the timed region; K steps move K batches).
        if graphed is None:
            last['loss'] = eager_step(to_dev(host[i % 2])).item()
            return
        loss = graphed.run_prefetched()
        graphed.prefetch(host[(i + 1) % 2])
        last['loss'] = loss.item()
    if graphed is not None: graphed.prefetch(host[0])
    for i in range(2): e2e_step(i)
    ms_e2e = timed(e2e_step, args.steps)

    if world > 1:
        tl = torch.tensor([launches], device=dev, dtype=torch.float64)
        dist.all_reduce(tl)
        launches = int(tl.item())

    if rank == 0:
        peak, peak_src = peaks()
        px = b*H*W
        per_px_fwd = 12 + 12*n_supp + 12*n_supp*N_SCALES + 4*N_SCALES    # target + identity supports + gathered supports + depth
        per_px_bwd = per_px_fwd + 4*N_SCALES                              # + d loss/d depth_up written
        bytes_fwd, bytes_bwd = per_px_fwd*px, per_px_bwd*px
        mean = lambda v: sum(v)/max(len(v), 1)
        t_f, t_b = mean(kt.get('stv_photo_fwd', [0])), mean(kt.get('stv_photo_bwd', [0]))
        gbs = lambda by, t: (by/1e9)/(t/1e3) if t > 0 else 0.0
        ach = gbs(bytes_fwd + bytes_bwd, t_f + t_b)
        traffic, traffic_src = photo_traffic() if args.config == 'c3' else (None, 'captured for configs[2] only')
        # Tensor-core kernel (every Linear / convolution product of the networks): algorithmic FLOPs from SURVEY 8d (forward GFLOP
        # per image x3 for forward + both gradients) over the summed live duration of the stv_gemm_tf32 / stv_conv_* calls of one
        # step; peak = cuBLAS TF32 measured on this GPU just now.
        tc_ms = sum(sum(kt.get(k, [])) for k in ('stv_gemm_tf32', 'stv_conv_fprop', 'stv_conv_dgrad', 'stv_conv_wgrad'))/TIMED_STEPS
        tc_flops = 3*c['gflop']*1e9*b
        tc_peak = measured_tf32_peak(dev)
        tc_ach = tc_flops/1e12/(tc_ms/1e3) if tc_ms > 0 else 0.0
        loss_ms = {k: round(mean(kt.get(k, [0])), 4) for k in ('stv_photo_fwd', 'stv_photo_bwd', 'stv_smooth_fwd', 'stv_smooth_bwd')}
        line = {
            'metric': 'training images/sec', 'value': round(b*world*args.steps/(ms/1e3), 3), 'unit': 'images/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': round(ms/args.steps, 3),
            'host_enqueue_ms_per_step': round(host_ms, 3), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'tf32', 'data': 'synthetic',
            'config': {'workload': c['label'], 'config': args.config, 'api': api, 'global_batch': b*world, 'per_gpu_batch': b,
                       'parallelism': f'dp{world}', 'launch': graph_note,
                       'params': n_params, 'optimizer': 'adamw(lr=1e-4, wd=1e-3), fused flat-buffer kernel per gradient bucket',
                       'l2_policy': f'inputs larger than L2: two rotating {h2d_bytes/1e6:.0f} MB batches + multi-GB activations per step',
                       'numerics': 'fp32 storage, TF32 tensor-core matmul/conv (reference: precision 32, matmul high), fp32 loss kernels'},
            'clocks': clocks,
            'e2e': {'value': round(b*world*args.steps/(ms_e2e/1e3), 3), 'unit': 'images/s', 'ms_per_step': round(ms_e2e/args.steps, 3),
                    'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4, 'last_loss': last.get('loss'), 'api': api},
            'gpu_launches': launches,
            'roofline': {'kernel': 'fused photometric loss: stv_photo_fused_fwd (loss + unit gradients in one sweep, three warps per strip) + stv_photo_fused_bwd',
                         'bound': 'hbm', 'achieved': round(ach, 1),
                         'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s', 'frac': round(ach/peak, 4),
                         'traffic': traffic, 'traffic_source': traffic_src,
                         'algorithmic_bytes_per_launch': bytes_fwd + bytes_bwd, 'algorithmic_bytes_per_pixel': per_px_fwd + per_px_bwd,
                         'avg_ms': round(t_f + t_b, 4),
                         'timing': f'CUDA events around the two entry points on their stream, mean of {TIMED_STEPS} eager single-stream steps of this configuration (a device-side spin in front of each step lets the host enqueue ahead, so no launch gap falls inside an event pair)',
                         'detail': {'photo_fwd_ms': loss_ms['stv_photo_fwd'], 'photo_bwd_ms': loss_ms['stv_photo_bwd'],
                                    'smooth_fwd_ms': loss_ms['stv_smooth_fwd'], 'smooth_bwd_ms': loss_ms['stv_smooth_bwd'],
                                    'loss_stack_ms': round(sum(loss_ms.values()), 4)}},
        }
        line['roofline_tensor'] = {
            'kernel': 'gemm_tf32_kernel (tcgen05 kind::tf32 + TMA/TMA-im2col + TMEM): all Linear / convolution fwd, dgrad, wgrad of the step',
            'bound': 'tensor', 'achieved': round(tc_ach, 1), 'peak': round(tc_peak, 1), 'unit': 'TFLOP/s', 'frac': round(tc_ach/tc_peak, 4),
            'peak_source': 'measured now: torch.matmul fp32 8192^3 with TF32 enabled (cuBLAS), best of 6', 'algorithmic_flops_per_step': tc_flops,
            'ms_per_step_in_kernel_calls': round(tc_ms, 3)}
        if use_plugin: plugin.uninstall()   # the baselines below run the reference's / the library's own classes
        if world == 1 and not args.no_torch_baseline:
            try: line['gpu_torch_baseline'] = torch_gpu_baseline(c, b, dev)
            except Exception as e: line['gpu_torch_baseline'] = {'unavailable': f'{type(e).__name__}: {str(e)[:160]}'}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference(c, steps=5, warmup=2, budget_s=60.0)
            line['cpu_baseline'] = {'value': round(r['value'], 4), 'unit': 'images/s', 'cores': r['cores'], 'kind': r['kind'], 'sample': r['sample']}
        print(json.dumps(line), file=out, flush=True)
    if world > 1:
        # The step graph holds captured NCCL collectives: release it before the communicator goes away, and never let a
        # shutdown problem hang the job after the result line is out (bounded wait, then a hard exit on every rank).
        import gc
        sync_all()
        graphed = train_step = None
        gc.collect()
        torch.cuda.synchronize()
        th = threading.Thread(target=dist.destroy_process_group, daemon=True)
        th.start()
        th.join(20.0)
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
