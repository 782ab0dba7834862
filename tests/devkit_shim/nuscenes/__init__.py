"""Minimal stand-in for the un-vendored nuscenes-devkit (requirements.txt:4 of the reference), TEST INFRASTRUCTURE ONLY: just the
table access the reference's loader uses (src/nuscenes_loader.py:7-8, 31, 66-99, 136-195) with the devkit's published semantics --
JSON tables under <dataroot>/<version>/, token lookup, the reverse indexing the devkit adds at load time (sample['data'],
sample['anns'], sample_data['channel'] / ['sensor_modality'], sample_annotation['category_name']), box_velocity and
LidarPointCloud.from_file.  It lets the real NuScenesLoader code paths (the reference's and this repo's) run against an on-disk tree
written by msc_geom.io.write_nuscenes_tree; it is not a devkit replacement."""
