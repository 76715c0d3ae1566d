"""`StatsAggregator`: the per-run statistics and the `run_data.csv` row of the reference (reference stats.py:12-199),
fed by batches of finished episodes instead of one episode dict at a time.

The reference keeps per-tick Python lists for one episode and reduces them when the episode ends
(stats.py:43-85).  Here thousands of episodes run in parallel on the device, so the per-tick reductions
(sums, minima, maxima, counts) are carried per episode by `control.EpisodeTracker` and only the finished
episodes' scalars arrive here.  Stat names, the `_std` columns (standard error of the mean, scipy.stats.sem),
ST_DESCRIPTION / TRAFFIC_DESCRIPTION and the scalar Settings columns are the reference's, so a row drops into the
reference's experiment notebook.  Plots are out of scope; the per-segment jerk / speed averages
the reference plots are kept as arrays and logged.
"""
from __future__ import annotations

import datetime
import logging
import os
from typing import Callable, Dict, Optional

import numpy as np

from .config import Settings

EPISODE_KEYS = ("crashed", "merged", "mean_speed", "max_speed", "mean_abs_jerk", "time_taken", "clock_time_per_episode",
                "clock_time_per_step")
OPTIONAL_KEYS = ("closest_distance", "mean_closest_distance", "mean_disruption", "max_disruption", "total_disruption",
                 "disruption_time")
MERGED_KEYS = ("time_to_merge", "closest_distance_merged", "mean_closest_distance_merged", "mean_abs_jerk_merged")
STAT_ORDER = ("crashed", "merged", "mean_speed", "max_speed", "mean_abs_jerk", "closest_distance", "mean_closest_distance",
              "mean_abs_jerk_merged", "closest_distance_merged", "mean_closest_distance_merged", "mean_disruption",
              "max_disruption", "total_disruption", "disruption_time", "time_taken", "time_to_merge",
              "clock_time_per_episode", "clock_time_per_step")


def sem(x) -> float:
    """Standard error of the mean with ddof=1 (scipy.stats.sem, the reference's `_std` columns)."""
    x = np.asarray(x, dtype=np.float64)
    return float(np.std(x, ddof=1) / np.sqrt(x.size)) if x.size > 1 else float("nan")


class StatsAggregator:
    def __init__(self, save_state_on_crash: bool = False):
        self.lists: Dict[str, list] = {k: [] for k in STAT_ORDER}
        self.custom_stat_lists: Dict[str, list] = {}
        self.bins = np.arange(-220, 61, 20)
        self.counts = np.zeros(len(self.bins) - 1)
        self.jerks = np.zeros(len(self.bins) - 1)
        self.speeds = np.zeros(len(self.bins) - 1)
        self.episodes = 0
        self.save_state_on_crash = save_state_on_crash       # accepted for signature parity; histories are not kept
        self.custom_stats_function: Optional[Callable[[dict], dict]] = None

    def add_custom_stat_callback(self, callback):
        self.custom_stats_function = callback

    # ---- one finished episode, already reduced (the reference's add_episode_stats, stats.py:43-85) ----
    def add_episode_stats(self, ep: dict):
        L = self.lists
        for k in EPISODE_KEYS:
            L[k].append(float(ep[k]))
        if ep.get("n_closest", 0) > 0:                                      # stats.py:58-60
            L["closest_distance"].append(float(ep["closest_distance"]))
            L["mean_closest_distance"].append(float(ep["mean_closest_distance"]))
        if ep.get("n_disruption", 0) > 0:                                   # stats.py:63-67
            for k in ("mean_disruption", "max_disruption", "total_disruption", "disruption_time"):
                L[k].append(float(ep[k]))
        if ep["merged"]:                                                    # stats.py:68-72
            L["time_to_merge"].append(float(ep["time_taken"]))
            L["mean_abs_jerk_merged"].append(float(ep["mean_abs_jerk"]))
            if ep.get("n_closest", 0) > 0:
                L["closest_distance_merged"].append(float(ep["closest_distance"]))
                L["mean_closest_distance_merged"].append(float(ep["mean_closest_distance"]))
        if self.custom_stats_function is not None:
            for key, value in self.custom_stats_function(ep).items():
                self.custom_stat_lists.setdefault(key, []).append(value)
        self.episodes += 1

    def add_segment_histograms(self, counts, jerks, speeds):
        """Per-20 m-segment tick counts, |jerk| sums and speed sums (stats.py:44-52), accumulated on the device."""
        self.counts += np.asarray(counts, dtype=np.float64)
        self.jerks += np.asarray(jerks, dtype=np.float64)
        self.speeds += np.asarray(speeds, dtype=np.float64)

    def get_stats(self):
        out = {k: self.lists[k] for k in STAT_ORDER}
        out.update(self.custom_stat_lists)
        return out

    def get_stat_averages(self, report_stds: bool = False):
        averages, stds = {}, {}
        for name, stat in self.get_stats().items():
            averages[name] = float(np.mean(stat)) if len(stat) else float("nan")
            stds[name] = sem(stat)
        return (averages, stds) if report_stds else averages

    def get_stat_report_row_dict(self):
        averages, stds = self.get_stat_averages(report_stds=True)
        columns = {}
        for name in averages:
            columns[name] = averages[name]
            columns[name + "_std"] = stds[name]
        for key, value in Settings.export_settings().items():
            if isinstance(value, (str, int, float)):
                columns[key] = value
        S = Settings
        if getattr(S, "USE_ALTERNATE_TRAFFIC_DISTRIBUTION", False):
            traffic = "joseph_{}".format(getattr(S, "TRAFFIC_DENSITY", ""))
        elif getattr(S, "USE_SIMPLE_TRAFFIC_DISTRIBUTION", True):
            traffic = "uniform-{}-{}-{}".format(S.OTHER_CAR_SPEED, S.BASE_TRAFFIC_INTERVAL,
                                                "varying" if S.VARY_TRAFFIC_START_TIMES else "constant")
        else:
            traffic = "harsh"
        columns["ST_DESCRIPTION"] = "st-{}-{}-{}-{}-{}-{}-{}-{}".format(                # stats.py:176-185 (A_WEIGHT twice, as there)
            S.V_WEIGHT, S.A_WEIGHT, S.J_WEIGHT, S.A_WEIGHT, S.MIN_ALLOWED_DISTANCE, S.CRASH_MIN_S, S.START_UNCERTAINTY,
            S.UNCERTAINTY_PER_SECOND)
        columns["TRAFFIC_DESCRIPTION"] = traffic
        columns["TIME"] = datetime.datetime.now().astimezone().isoformat()
        return columns

    def print_stats(self, csv_path: str = "run_data.csv"):
        with np.errstate(invalid="ignore", divide="ignore"):
            avg_jerks, avg_speeds = self.jerks / self.counts, self.speeds / self.counts
        logging.info("Average jerks per segment: %s", dict(zip(self.bins[:-1].tolist(), np.round(avg_jerks, 4).tolist())))
        logging.info("Average speeds per segment: %s", dict(zip(self.bins[:-1].tolist(), np.round(avg_speeds, 4).tolist())))
        averages, stds = self.get_stat_averages(report_stds=True)
        for name in averages:
            message = "{}: {} ± {}".format(name, averages[name], stds[name])
            logging.info(message)
            print(message)
        self.add_csv_data(csv_path)

    def add_csv_data(self, csv_path: str = "run_data.csv"):
        """Append this run as one row (reference stats.py:192-199: read, concat, rewrite)."""
        import pandas as pd
        new_row = pd.DataFrame([self.get_stat_report_row_dict()])
        if os.path.exists(csv_path):
            new_row = pd.concat([pd.read_csv(csv_path), new_row], ignore_index=True)
        new_row.to_csv(csv_path, index=False)
