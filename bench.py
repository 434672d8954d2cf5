#!/usr/bin/env python
"""Benchmark of the SlowTV-monodepth training hot path on B200 (one JSON line on stdout from rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one full training step of BASELINE.json config 3 (ConvNeXt-T depth + ResNet-18 pose, 384x640, batch 8 per GPU,
2 support frames, 4 scales, min-reprojection + automask + edge-aware smoothness, AdamW): networks forward, fused loss,
backward, gradient all-reduce (N > 1), optimiser step. Synthetic video triplets, random-init weights.

  value  images/s over all ranks, inputs resident in HBM before the timed region (CUDA events, max over ranks)
  e2e    the same metric through the public API with HOST (pinned) batches: H2D copy of every step's batch and a D2H
         read of the loss inside the timed region
  roofline      fused photometric loss (forward + backward entry points), algorithmic bytes (148 + 164 B per
                target pixel for n=2, S=4) over the live CUDA-event duration of those calls, vs MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the reference step (oracle/step.py, PyTorch CPU, all host threads) on a bounded sample

`--impl reference` times that CPU port alone (the reference is pure Python/PyTorch and its tree is not present on the GPU
box, so the arm executes the oracle restatement of it; rank 0 only).
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

SHAPE, BATCH, N_SUPP, N_SCALES = (384, 640), 8, 2, 4
DEPTH_ENC, POSE_ENC = 'convnext_tiny', 'resnet18'
WORKLOAD = 'configs[2]: ConvNeXt-T depth + ResNet-18 pose (KBR default), 384x640, batch 8/GPU, 2 support frames, S=4'
CPU_BATCH = 2


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=BATCH)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='enqueue every step eagerly instead of replaying the captured CUDA graph')
    return ap.parse_args()


def peaks() -> tuple[float, str]:
    f = ROOT/'MEASURED_PEAKS.json'
    if f.is_file():
        return float(json.loads(f.read_text())['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: oracle port of the reference step
# ---------------------------------------------------------------------------------------------------------------------
def reference_cfg(batch: int) -> dict:
    """The reference's own experiment configuration of the benchmarked workload: cfg/abl_learn_K/default.yaml (ConvNeXt-T depth +
    ResNet-18 pose) with the KBR loss / depth-range / optimiser settings (cfg/kbr/default.yaml), random init."""
    from oracle import ref_shim
    cfg = ref_shim.load_cfg('abl_learn_K/default.yaml')
    kbr = ref_shim.load_cfg('kbr/default.yaml')
    cfg['net']['depth'].update(enc_name=DEPTH_ENC, pretrained=False)
    cfg['net']['pose'].update(enc_name=POSE_ENC, pretrained=False, learn_K=False)
    cfg['loss'] = {k: kbr['loss'][k] for k in ('img_recon', 'disp_smooth')}
    cfg['optimizer'] = {'type': 'adamw', 'lr': 1e-4, 'weight_decay': 1e-3}
    cfg['scheduler'] = None
    cfg['loader'] = {'batch_size': batch}
    cfg['dataset'] = {}
    cfg['trainer'].update(min_depth=0.1, max_depth=100, always_fwd_pose=False, aspect_ratio_aug_prob=0.0)
    return cfg


def cpu_reference(steps: int, warmup: int, batch: int = CPU_BATCH, budget_s: float = 200.0) -> dict:
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
    batches = [syn.make_batch(batch, N_SUPP, SHAPE, seed=s) for s in range(2)]
    if ref_shim.available():
        warnings.filterwarnings('ignore')
        ref_shim.load()
        import src.core.trainer as rt
        from src.tools import parsers
        cfg = reference_cfg(batch)
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
        tr = OracleTrainer(DEPTH_ENC, POSE_ENC).train()
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
    r = cpu_reference(steps, warmup)
    line = {
        'impl': 'reference', 'metric': 'training images/sec', 'value': round(r['value'], 4), 'unit': 'images/s',
        'n_gpus': args.gpus, 'steps': r['steps_done'], 'warmup': warmup, 'ms_per_step': round(r['ms_per_step'], 2),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'cpu_batch_per_step': r['batch'], 'requested_steps': args.steps,
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


def main() -> None:
    args = parse()
    out = _claim_stdout()
    if args.impl == 'reference': return run_reference_arm(args, out)

    import torch.distributed as dist
    from slowtv_monodepth_b200 import _lib as L, functional as F_, synthetic as syn
    from slowtv_monodepth_b200.optim import FlatAdamW
    from slowtv_monodepth_b200.trainer import GraphedTrainStep, MonoDepthStep, default_cfg

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
    b, (H, W) = args.batch, SHAPE
    torch.manual_seed(1234)  # identical initial weights on every rank (DDP broadcast equivalent)
    model = MonoDepthStep(default_cfg(DEPTH_ENC, POSE_ENC)).to(dev).train()
    model = model.to(memory_format=torch.channels_last)
    opt = FlatAdamW(model.nets, lr=1e-4, weight_decay=1e-3)
    n_params = opt.flat.numel()

    # Distinct batches per rank (DistributedSampler equivalent): seed = base + rank. Two batches are rotated.
    host = [syn.make_batch(b, N_SUPP, SHAPE, seed=100*rank + s, pin=True) for s in range(2)]
    to_dev = lambda bt: ({k: (v.to(dev, non_blocking=True) if k != 'supp_idxs' else v) for k, v in bt[0].items()},
                         {k: v.to(dev, non_blocking=True) for k, v in bt[1].items()}, {})
    resident = [to_dev(bt) for bt in host]
    h2d_bytes = sum(v.numel()*v.element_size() for d in host[0][:2] for k, v in d.items() if k != 'supp_idxs')

    def eager_step(batch):
        opt.zero_grad()
        loss, _, _ = model.step(batch)
        loss.backward()
        opt.all_reduce_async()
        opt.step()
        return loss

    for i in range(3): eager_step(resident[i % 2])   # first-use initialisation; also the photometric kernels' live timing below
    F_.enable_kernel_timing(True)      # CUDA events around each libstv call, on the launching stream
    eager_step(resident[1])
    torch.cuda.synchronize()
    F_.reset_kernel_timings()
    TIMED_STEPS = 3
    for i in range(TIMED_STEPS): eager_step(resident[i % 2])
    torch.cuda.synchronize()
    kt = F_.kernel_timings()
    F_.enable_kernel_timing(False)
    graphed, graph_note = None, 'eager (--no-graph)'
    if not args.no_graph:
        try:
            graphed = GraphedTrainStep(model, opt, resident[0])
            graph_note = 'CUDA graph replay of zero_grad+fwd+loss+bwd, then all-reduce + AdamW'
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
        bytes_fwd, bytes_bwd = (12 + 12*N_SUPP + 12*N_SUPP*N_SCALES + 4*N_SCALES)*px, (12 + 12*N_SUPP + 12*N_SUPP*N_SCALES + 8*N_SCALES)*px
        mean = lambda v: sum(v)/max(len(v), 1)
        t_f, t_b = mean(kt.get('stv_photo_fwd', [0])), mean(kt.get('stv_photo_bwd', [0]))
        gbs = lambda by, t: (by/1e9)/(t/1e3) if t > 0 else 0.0
        ach = gbs(bytes_fwd + bytes_bwd, t_f + t_b)
        # Tensor-core kernel (every Linear / convolution product of the networks): algorithmic FLOPs of config 3 from SURVEY 8d
        # (forward 43.6 + 14.0 + 2 x 20.7 GFLOP per image, x3 for forward + both gradients) over the summed live duration of
        # the stv_gemm_tf32 / stv_conv_* calls of one step; peak = TF32 dense = half the measured bf16 figure.
        tc_ms = sum(sum(kt.get(k, [])) for k in ('stv_gemm_tf32', 'stv_conv_fprop', 'stv_conv_dgrad', 'stv_conv_wgrad'))/TIMED_STEPS
        tc_flops = 3*(43.6 + 14.0 + N_SUPP*20.7)*1e9*b*(H*W)/(384*640)
        pk = json.loads((ROOT/'MEASURED_PEAKS.json').read_text()) if (ROOT/'MEASURED_PEAKS.json').is_file() else {}
        tc_peak = float(pk.get('bf16_tflops_sustained', 1400.0))/2
        tc_ach = tc_flops/1e12/(tc_ms/1e3) if tc_ms > 0 else 0.0
        line = {
            'metric': 'training images/sec', 'value': round(b*world*args.steps/(ms/1e3), 3), 'unit': 'images/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': round(ms/args.steps, 3),
            'host_enqueue_ms_per_step': round(host_ms, 3), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'tf32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'global_batch': b*world, 'per_gpu_batch': b, 'parallelism': f'dp{world}', 'launch': graph_note,
                       'params': n_params, 'optimizer': 'adamw(lr=1e-4, wd=1e-3), fused flat-buffer kernel',
                       'l2_policy': 'inputs larger than L2: two rotating 141 MB batches + multi-GB activations per step',
                       'numerics': 'fp32 storage, TF32 tensor-core matmul/conv (reference: precision 32, matmul high), fp32 loss kernels'},
            'clocks': clocks,
            'e2e': {'value': round(b*world*args.steps/(ms_e2e/1e3), 3), 'unit': 'images/s', 'ms_per_step': round(ms_e2e/args.steps, 3),
                    'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4, 'last_loss': last.get('loss')},
            'gpu_launches': launches,
            'roofline': {'kernel': 'fused photometric loss, stv_photo_fwd + stv_photo_bwd', 'bound': 'hbm', 'achieved': round(ach, 1),
                         'peak': peak, 'peak_source': peak_src, 'unit': 'GB/s', 'frac': round(ach/peak, 4),
                         'traffic': 1252.0e6, 'traffic_source': 'ncu --set full dram__bytes_read+write: fwd 343 + 297 MB, bwd 579 + 33 MB (profiles/r1f_loss_ncu_full_summary.txt); the 2x over the algorithmic bytes is the (S,b,9,H,W) SSIM coefficient planes the forward hands to the backward (283 MB written, re-read with a 1-pixel halo) — traded for a 2.4x shorter backward',
                         'algorithmic_bytes_per_launch': bytes_fwd + bytes_bwd, 'avg_ms': round(t_f + t_b, 4),
                         'detail': {'photo_fwd': {'ms': round(t_f, 4), 'GB/s': round(gbs(bytes_fwd, t_f), 1), 'frac': round(gbs(bytes_fwd, t_f)/peak, 4)},
                                    'photo_bwd': {'ms': round(t_b, 4), 'GB/s': round(gbs(bytes_bwd, t_b), 1), 'frac': round(gbs(bytes_bwd, t_b)/peak, 4)},
                                    'smooth_fwd_ms': round(mean(kt.get('stv_smooth_fwd', [0])), 4),
                                    'smooth_bwd_ms': round(mean(kt.get('stv_smooth_bwd', [0])), 4)}},
        }
        line['roofline_tensor'] = {
            'kernel': 'gemm_tf32_kernel (tcgen05 kind::tf32 + TMA/TMA-im2col + TMEM): all Linear / convolution fwd, dgrad, wgrad of the step',
            'bound': 'tensor', 'achieved': round(tc_ach, 1), 'peak': tc_peak, 'unit': 'TFLOP/s', 'frac': round(tc_ach/tc_peak, 4),
            'peak_source': 'TF32 dense = MEASURED_PEAKS.json bf16_tflops_sustained / 2', 'algorithmic_flops_per_step': tc_flops,
            'ms_per_step_in_kernel_calls': round(tc_ms, 3)}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference(steps=5, warmup=2, budget_s=60.0)
            line['cpu_baseline'] = {'value': round(r['value'], 4), 'unit': 'images/s', 'cores': r['cores'], 'kind': r['kind'], 'sample': r['sample']}
        print(json.dumps(line), file=out, flush=True)
    if world > 1: dist.destroy_process_group()


if __name__ == '__main__':
    main()
