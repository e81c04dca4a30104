"""Speed benchmark with the reference's result schema.

``benchmark_speed`` mirrors experiments/benchmark.py:30-81 of the reference -- same arguments, same timing loop
(warm-up sample on the first run, ``torch.cuda.synchronize`` around ``time.time``, mean / std over ``num_runs``),
same dictionary keys per step count -- with the per-call batch size exposed as ``batch_size`` (the reference
hard-codes 4, experiments/benchmark.py:60-61).  ``write_results_csv`` writes the file the reference writes at
experiments/benchmark.py:252-262: columns ``num_steps,base_time_ms,rect_time_ms,base_img_per_sec,rect_img_per_sec,
speedup`` so results drop into its ``results/`` layout and plotting code; ``create_summary_report`` writes the
``benchmark_report.txt`` of utils/visualization.py:210-253 next to it.

    python -m rectified_flow_vision_b200.benchmark --out results/benchmark_results.csv [--base ckpt.pt --rect ckpt.pt]
"""
from __future__ import annotations

import argparse
import csv
import os
import time
from typing import Dict, List

import numpy as np
import torch


def benchmark_speed(model, num_samples: int, steps_list: list, image_size: int, device: str, num_runs: int = 3,
                    batch_size: int = 4) -> List[Dict]:
    model.eval()
    results = []
    for num_steps in steps_list:
        times = []
        for run in range(num_runs):
            if run == 0:  # warm-up on the first run
                noise = torch.randn(1, 3, image_size, image_size, device=device)
                with torch.no_grad():
                    _ = model.sample(noise=noise, num_steps=num_steps)
            if str(device).startswith('cuda'):
                torch.cuda.synchronize()
            start_time = time.time()
            with torch.no_grad():
                for i in range(0, num_samples, batch_size):
                    bs = min(batch_size, num_samples - i)
                    noise = torch.randn(bs, 3, image_size, image_size, device=device)
                    _ = model.sample(noise=noise, num_steps=num_steps)
            if str(device).startswith('cuda'):
                torch.cuda.synchronize()
            times.append(time.time() - start_time)
        avg_time, std_time = float(np.mean(times)), float(np.std(times))
        results.append({'num_steps': num_steps, 'total_time': avg_time, 'time_per_image': avg_time / num_samples,
                        'images_per_second': num_samples / avg_time, 'time_std': std_time, 'num_samples': num_samples})
    return results


def results_table(base_results: List[Dict], rect_results: List[Dict]) -> List[Dict]:
    rows = []
    for b, r in zip(base_results, rect_results):
        bt, rt = b['time_per_image'] * 1000, r['time_per_image'] * 1000
        rows.append({'num_steps': b['num_steps'], 'base_time_ms': bt, 'rect_time_ms': rt,
                     'base_img_per_sec': b['images_per_second'], 'rect_img_per_sec': r['images_per_second'],
                     'speedup': bt / rt})
    return rows


def write_results_csv(path: str, base_results: List[Dict], rect_results: List[Dict]) -> None:
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    rows = results_table(base_results, rect_results)
    with open(path, 'w', newline='') as f:
        w = csv.DictWriter(f, fieldnames=['num_steps', 'base_time_ms', 'rect_time_ms', 'base_img_per_sec',
                                          'rect_img_per_sec', 'speedup'])
        w.writeheader()
        w.writerows(rows)


def summary_report_text(results: Dict) -> str:
    """The text of ``benchmark_report.txt`` as the reference lays it out (utils/visualization.py:219-251): a speed table with
    one row per step count (ms per image for both models and their ratio) and the mean / max / min speed-up.  ``results`` is
    the reference's dictionary: ``{'base_model': [...], 'rectified_model': [...]}`` of ``benchmark_speed`` rows."""
    rule, thin = "=" * 60, "-" * 40
    pairs = list(zip(results['base_model'], results['rectified_model']))
    out = [rule, "BENCHMARK REPORT: FLOW DISTILLATION", rule, "", "SPEED COMPARISON", thin,
           f"{'Steps':<10} {'Base (ms/img)':<15} {'Rect (ms/img)':<15} {'Speedup':<10}", thin]
    ratios = []
    for b, r in pairs:
        bt, rt = b['time_per_image'] * 1000, r['time_per_image'] * 1000
        out.append(f"{b['num_steps']:<10} {bt:<15.2f} {rt:<15.2f} {(bt / rt if rt > 0 else 0):<10.2f}x")
        if r['time_per_image'] > 0:
            ratios.append(b['time_per_image'] / r['time_per_image'])
    out += ["", rule, "CONCLUSIONS", thin]
    if ratios:
        out += [f"Average speedup: {float(np.mean(ratios)):.2f}x", f"Maximum speedup: {max(ratios):.2f}x",
                f"Minimum speedup: {min(ratios):.2f}x"]
    return "\n".join(out) + "\n"


def create_summary_report(results: Dict, save_dir: str) -> str:
    """Writes ``<save_dir>/benchmark_report.txt`` (utils/visualization.py:210-253, the text half; the matplotlib plot of that
    function is out of scope) and returns its path."""
    os.makedirs(save_dir, exist_ok=True)
    path = os.path.join(save_dir, 'benchmark_report.txt')
    with open(path, 'w') as f:
        f.write(summary_report_text(results))
    print(f"Report saved to: {path}")
    return path


def main():
    from . import BaseFlowModel, RectifiedFlowModel
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default='results/benchmark_results.csv')
    ap.add_argument('--base', default=None, help='base_flow_final.pt (default: seeded random-init weights)')
    ap.add_argument('--rect', default=None, help='rectified_flow_k1_final.pt')
    ap.add_argument('--image-size', type=int, default=64)
    ap.add_argument('--num-samples', type=int, default=64)
    ap.add_argument('--batch-size', type=int, default=64)
    ap.add_argument('--steps', type=int, nargs='+', default=[1, 2, 4, 8, 16, 32, 64, 100])
    a = ap.parse_args()
    dev = 'cuda'
    torch.manual_seed(0)
    base = BaseFlowModel(image_size=a.image_size, device=dev)
    rect = RectifiedFlowModel(image_size=a.image_size, device=dev)
    if a.base:
        base.load(a.base)
    if a.rect:
        rect.load(a.rect)
    br = benchmark_speed(base, a.num_samples, a.steps, a.image_size, dev, batch_size=a.batch_size)
    rr = benchmark_speed(rect, a.num_samples, a.steps, a.image_size, dev, batch_size=a.batch_size)
    write_results_csv(a.out, br, rr)
    create_summary_report({'base_model': br, 'rectified_model': rr}, os.path.dirname(a.out) or '.')
    for row in results_table(br, rr):
        print(row)


if __name__ == '__main__':
    main()
