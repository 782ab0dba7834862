import numpy as np


class LidarPointCloud:
    """points: (4, N) float32 -- the devkit drops the fifth (ring) float of every .pcd.bin row."""

    def __init__(self, points: np.ndarray):
        self.points = points

    @staticmethod
    def nbr_dims() -> int:
        return 4

    @classmethod
    def from_file(cls, file_name: str) -> "LidarPointCloud":
        assert file_name.endswith(".bin"), "Unsupported filetype {}".format(file_name)
        scan = np.fromfile(file_name, dtype=np.float32)
        points = scan.reshape((-1, 5))[:, :cls.nbr_dims()]
        return cls(points.T)
