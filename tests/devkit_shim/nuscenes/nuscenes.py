import json
import os

import numpy as np

_TABLES = ("category", "attribute", "visibility", "instance", "sensor", "calibrated_sensor", "ego_pose", "log", "scene", "sample",
           "sample_data", "sample_annotation", "map")


class NuScenes:
    def __init__(self, version: str = "v1.0-mini", dataroot: str = "/data/sets/nuscenes", verbose: bool = True, map_resolution: float = 0.1):
        self.version, self.dataroot, self.verbose = version, dataroot, verbose
        self.table_root = os.path.join(dataroot, version)
        self._token2ind = {}
        for name in _TABLES:
            with open(os.path.join(self.table_root, name + ".json")) as f:
                rows = json.load(f)
            setattr(self, name, rows)
            self._token2ind[name] = {r["token"]: i for i, r in enumerate(rows)}
        # reverse indexing, as the devkit's __make_reverse_index__ does
        for rec in self.sample_annotation:
            inst = self.get("instance", rec["instance_token"])
            rec["category_name"] = self.get("category", inst["category_token"])["name"]
        for rec in self.sample_data:
            cs = self.get("calibrated_sensor", rec["calibrated_sensor_token"])
            sensor = self.get("sensor", cs["sensor_token"])
            rec["sensor_modality"], rec["channel"] = sensor["modality"], sensor["channel"]
        for rec in self.sample:
            rec["data"], rec["anns"] = {}, []
        for rec in self.sample_data:
            if rec["is_key_frame"]:
                self.get("sample", rec["sample_token"])["data"][rec["channel"]] = rec["token"]
        for rec in self.sample_annotation:
            self.get("sample", rec["sample_token"])["anns"].append(rec["token"])

    def get(self, table_name: str, token: str) -> dict:
        return getattr(self, table_name)[self._token2ind[table_name][token]]

    def box_velocity(self, sample_annotation_token: str, max_time_diff: float = 1.5) -> np.ndarray:
        """Devkit semantics: finite difference of the annotation's centre between its prev and next records (falling back to the record
        itself on either side); NaN when there is no neighbour or the time gap exceeds max_time_diff."""
        current = self.get("sample_annotation", sample_annotation_token)
        has_prev, has_next = current["prev"] != "", current["next"] != ""
        if not has_prev and not has_next:
            return np.array([np.nan, np.nan, np.nan])
        first = self.get("sample_annotation", current["prev"]) if has_prev else current
        last = self.get("sample_annotation", current["next"]) if has_next else current
        t_last = 1e-6 * self.get("sample", last["sample_token"])["timestamp"]
        t_first = 1e-6 * self.get("sample", first["sample_token"])["timestamp"]
        dt = t_last - t_first
        if has_next and has_prev:
            max_time_diff *= 2
        if dt > max_time_diff:
            return np.array([np.nan, np.nan, np.nan])
        return (np.array(last["translation"]) - np.array(first["translation"])) / dt
