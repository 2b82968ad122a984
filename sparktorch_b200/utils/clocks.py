"""Sample GPU clocks / throttle reasons while a benchmark runs
(B200_PROFILING.md 'clocks DURING the timed region')."""
from __future__ import annotations

import statistics
import subprocess
import threading
from typing import Dict, List, Optional

_Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
      "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
      "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler:
    def __init__(self, period_ms: int = 100, gpus: Optional[List[int]] = None):
        self.period_ms = period_ms
        self.gpus = gpus
        self.proc = None
        self.lines: List[str] = []
        self.stamps: List[float] = []
        self._thr = None

    def start(self) -> "ClockSampler":
        cmd = ["nvidia-smi", "--query-gpu=" + _Q, "--format=csv,noheader,nounits", "-lms", str(self.period_ms)]
        if self.gpus is not None:
            cmd += ["-i", ",".join(str(g) for g in self.gpus)]
        try:
            self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except (FileNotFoundError, OSError):
            self.proc = None
            return self

        def pump():
            import time

            for line in self.proc.stdout:  # type: ignore[union-attr]
                self.lines.append(line.strip())
                self.stamps.append(time.time())

        self._thr = threading.Thread(target=pump, daemon=True)
        self._thr.start()
        return self

    def mark(self) -> int:
        return len(self.lines)

    def stop(self) -> None:
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=3)
            except Exception:
                self.proc.kill()

    def window(self, t0: float, t1: float, margin: float = 0.12):
        """Indices [lo, hi) of the samples taken between host times t0 and t1 (+- margin seconds)."""
        idx = [i for i, t in enumerate(self.stamps) if t0 - margin <= t <= t1 + margin]
        if not idx:
            return 0, 0
        return idx[0], idx[-1] + 1

    def summary(self, lo: int = 0, hi: Optional[int] = None) -> Dict:
        rows = [ln.split(", ") for ln in self.lines[lo:hi] if ln.count(",") >= 8]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [float(r[1]) for r in rows]
        smax = max(float(r[2]) for r in rows)
        power = [float(r[3]) for r in rows if r[3].replace(".", "", 1).isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[5 + k].strip().lower() == "active" for r in rows)]
        return {"sm_mhz": statistics.median(sm), "sm_min_mhz": min(sm), "sm_max_mhz": smax,
                "power_w_max": max(power) if power else None, "reasons": reasons, "samples": len(rows)}
