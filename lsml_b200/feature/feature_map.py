"""FeatureMap -- host-side mirror of lsml/feature/feature_map.py on top of the device engine."""
import numpy as np

from .base_feature import BaseFeature, BaseImageFeature, BaseShapeFeature


def feature_specs(features):
    return [f.spec() for f in features]


def compute_dense(features, u, img, mask, dx):
    """Run the device feature program for one item with a caller-supplied band mask and
    scatter the (B, F) rows into a dense ``u.shape + (F,)`` float64 array (zero outside the
    mask, as feature_map.py:62-63)."""
    import torch
    from ..core.engine import BandEngine
    specs = feature_specs(features)
    eng = BandEngine(u.shape, 1, dx=dx, band=0.0, specs=specs)
    if img is None:
        img = np.zeros(u.shape)
    eng.load_images(np.ascontiguousarray(img, dtype=np.float64)[None], normalize=False)
    eng.set_level_sets(np.ascontiguousarray(u, dtype=np.float64)[None], mask=mask[None])
    X = eng.compute_features()
    nfeat = eng.program.nfeat
    out = torch.zeros((u.size, nfeat), dtype=torch.float64, device=eng.device)
    if eng.n_band:
        out[eng.band_idx[:eng.n_band]] = X
    return out.cpu().numpy().reshape(u.shape + (nfeat,))


class FeatureMap(object):
    """ Stores a set of features and computes their results into a stacked array
    (reference feature_map.py:7-80) """

    def __init__(self, features):
        self._validate_features(features)
        self.features = features

    @property
    def n_features(self):
        return sum([f.size for f in self.features])

    def _validate_features(self, features):
        if not np.iterable(features):
            raise TypeError("features was not iterable")
        for feature in features:
            if not isinstance(feature, BaseFeature):
                msg = "feature {} was not an instance of {}".format(
                    feature, BaseFeature.__class__.__name__)
                raise ValueError(msg)
        if len(set(features)) != len(features):
            raise ValueError("`features` list included non-unique features")

    @property
    def feature_slices(self):
        j = 0
        indices = []
        for feature in self.features:
            if feature.size > 1:
                indices.append(slice(j, j + feature.size))
            else:
                indices.append(j)
            j += feature.size
        return indices

    def specs(self):
        for feature in self.features:
            if not isinstance(feature, (BaseImageFeature, BaseShapeFeature)):
                raise ValueError("Unknown feature type ({})".format(feature.__class__.__name__))
        return feature_specs(self.features)

    def __call__(self, u, img, dist, mask, dx=None):
        """ features: numpy.array, shape = img.shape + (n_features,) (reference :53-80) """
        if not mask.any():
            return np.zeros(u.shape + (self.n_features,))
        self.specs()
        dx = np.ones(u.ndim) if dx is None else np.asarray(dx, dtype=float)
        return compute_dense(self.features, u, img, mask, dx)
