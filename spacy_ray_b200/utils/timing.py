"""Device-side timing + clock sampling (the reference has only an
unused wall-clock ``Timer``).  Every throughput number is measured with CUDA
events on the launching stream; ``ClockSampler`` records SM clocks / throttle
reasons during the timed region (B200_PROFILING.md recipe)."""
from __future__ import annotations

import subprocess
import threading
import time
from typing import Any, Dict, List, Optional

import torch


class CudaStepTimer:
    """Start/stop CUDA events around a region; ``elapsed_ms`` synchronises."""

    def __init__(self, device: Optional[torch.device] = None):
        self.cuda = torch.cuda.is_available() and (device is None or torch.device(device).type == "cuda")
        if self.cuda:
            self._start = torch.cuda.Event(enable_timing=True)
            self._stop = torch.cuda.Event(enable_timing=True)
        self._t0 = 0.0
        self._t1 = 0.0

    def start(self) -> None:
        if self.cuda:
            self._start.record()
        else:
            self._t0 = time.perf_counter()

    def stop(self) -> None:
        if self.cuda:
            self._stop.record()
        else:
            self._t1 = time.perf_counter()

    def elapsed_ms(self) -> float:
        if self.cuda:
            self._stop.synchronize()
            return float(self._start.elapsed_time(self._stop))
        return (self._t1 - self._t0) * 1e3


class ClockSampler:
    """Samples SM clock, power and throttle reasons in a background thread while a region is timed.

    NVML (``pynvml``) is polled every ``period_s`` (default 2 ms: a 20-step timed region is ~30 ms,
    one ``nvidia-smi`` process launch takes longer than that); ``nvidia-smi`` is the fallback when
    NVML cannot be loaded."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    # nvmlClocksEventReason* bits
    _BITS = (("hw_slowdown", 0x8), ("hw_thermal", 0x40), ("sw_thermal", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index: int = 0, period_s: float = 0.002):
        self.gpu_index = gpu_index
        self.period_s = period_s
        self.samples: List[Dict[str, Any]] = []
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self._nvml = None
        self._handle = None
        self.source = "nvidia-smi"
        try:
            import pynvml

            pynvml.nvmlInit()
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(gpu_index))
            self._sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
            self.source = "nvml"
        except Exception:
            self._nvml = None

    @staticmethod
    def _physical_index(i: int) -> int:
        import os

        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v.strip() for v in vis.split(",") if v.strip()]
        if ids and all(v.isdigit() for v in ids) and i < len(ids):
            return int(ids[i])
        return i

    def _sample_nvml(self) -> None:
        n = self._nvml
        h = self._handle
        try:
            reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = int(reasons_fn(h))
        except Exception:
            bits = 0
        s = {"sm": float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)), "sm_max": self._sm_max,
             "power": n.nvmlDeviceGetPowerUsage(h) / 1000.0, "active": hex(bits)}
        for key, bit in self._BITS:
            s[key] = "Active" if bits & bit else "Not Active"
        self.samples.append(s)

    def _sample_smi(self) -> None:
        out = subprocess.run(
            ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i", str(self.gpu_index)],
            capture_output=True, text=True, timeout=5,
        ).stdout.strip().splitlines()
        for line in out:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 9:
                self.samples.append({
                    "sm": float(f[1]), "sm_max": float(f[2]), "power": float(f[3]) if f[3] not in ("[N/A]", "N/A") else 0.0,
                    "active": f[4], "hw_slowdown": f[5], "hw_thermal": f[6], "sw_thermal": f[7], "sw_power_cap": f[8],
                })

    def _poll(self) -> None:
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(self.period_s if self._nvml is not None else max(self.period_s, 0.05))

    def __enter__(self) -> "ClockSampler":
        self._thread = threading.Thread(target=self._poll, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc) -> None:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=6)

    def summary(self) -> Dict[str, Any]:
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": self.source}
        sms = sorted(s["sm"] for s in self.samples)
        reasons = set()
        for s in self.samples:
            for key, name in (("hw_slowdown", "hw_slowdown"), ("hw_thermal", "hw_thermal_slowdown"),
                              ("sw_thermal", "sw_thermal_slowdown"), ("sw_power_cap", "sw_power_cap")):
                if str(s[key]).lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": sms[len(sms) // 2], "sm_min_mhz": sms[0], "sm_max_mhz": self.samples[0]["sm_max"],
            "power_w_max": max(s["power"] for s in self.samples), "reasons": sorted(reasons),
            "samples": len(self.samples), "source": self.source,
        }
