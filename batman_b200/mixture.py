"""Mixture of experts with device Kriging local models (surrogate/mixture.py:52-353).

Unsupervised learning splits the DoE into clusters, a classifier sends new
points to a cluster, and one local surrogate per cluster predicts them.  The
clusterer / classifier are the same scikit-learn objects the reference builds
(mixture.py:126-169); the local models are ``batman_b200.SurrogateModel``
(kind 'kriging'), so every prediction is one batched device call per cluster.
Not restated: ``boundaries`` (plotting) and local PODs (``pod=``).
"""
import logging

import numpy as np
import sklearn.cluster          # noqa: F401  (names resolved by the eval below, as in the reference)
import sklearn.ensemble         # noqa: F401
import sklearn.gaussian_process  # noqa: F401
import sklearn.mixture          # noqa: F401
import sklearn.naive_bayes      # noqa: F401
import sklearn.neighbors        # noqa: F401
import sklearn.svm              # noqa: F401
from sklearn import preprocessing
from sklearn.decomposition import PCA

from .surrogate_model import SurrogateModel


def _sklearn_object(spec, what):
    """mixture.py:126-142, 153-169: an estimator instance, or a string evaluated against the sklearn namespace."""
    if hasattr(spec, 'get_params'):
        return spec
    try:
        return eval("sklearn." + spec, {'__builtins__': None}, {'sklearn': sklearn})
    except (TypeError, AttributeError, NameError, SyntaxError):
        raise AttributeError('%s unknown from sklearn.' % what)


class Mixture:
    """Mixture class (arguments as mixture.py:61-99)."""

    logger = logging.getLogger(__name__)

    def __init__(self, samples, data, corners, fsizes=None, pod=None, standard=True, local_method=None,
                 pca_percentage=0.8, clusterer='cluster.KMeans(n_clusters=2)',
                 classifier='gaussian_process.GaussianProcessClassifier()'):
        if pod is not None:
            raise NotImplementedError("Mixture(pod=...): local PODs are fitted by batman.pod.Pod, outside this path")
        samples = np.asarray(samples, dtype=np.float64)
        data = np.asarray(data, dtype=np.float64)
        self.scaler = preprocessing.MinMaxScaler()
        self.scaler.fit(np.array(corners))
        samples = self.scaler.transform(samples)
        corners = [[0 for _ in range(samples.shape[1])], [1 for _ in range(samples.shape[1])]]
        self.fsizes = data.shape[1] if fsizes is None else fsizes
        clust = data[:, self.fsizes:] if data.shape[1] > self.fsizes else data       # mixture.py:110-113
        if clust.shape[1] > 1:                                                        # mixture.py:115-118
            pca = PCA(n_components=pca_percentage)
            pca.fit(clust.T)
            clust = pca.components_.T
        if standard is True:
            clust = preprocessing.StandardScaler().fit_transform(clust)
        clusterer = _sklearn_object(clusterer, 'Clusterer')
        try:
            labels = clusterer.fit_predict(clust)
        except AttributeError:
            clusterer.fit(clust)
            labels = clusterer.predict(clust)
        self.clusters_id = np.unique(labels)
        self.classifier = _sklearn_object(classifier, 'Classifier')
        self.classifier.fit(samples, labels)
        self.indice_clt, self.local_models = {}, {}
        for i, k in enumerate(self.clusters_id):
            idx = np.where(labels == k)[0]
            self.indice_clt[k] = list(idx)
            if local_method is None:
                self.local_models[k] = SurrogateModel('kriging', corners, plabels=None)
            else:
                method = [lm for lm in local_method[i]][0]
                self.local_models[k] = SurrogateModel(method, corners, plabels=None, **local_method[i][method])
            self.local_models[k].fit(samples[idx], data[idx, :self.fsizes])

    def estimate_quality(self, multi_q2=False):
        """mixture.py:279-303."""
        q2, point = {}, {}
        for k in self.clusters_id:
            q2[k], point[k] = self.local_models[k].estimate_quality()
        if multi_q2:
            return q2, point
        ind = min(q2, key=q2.get)
        return q2[ind], point[ind]

    def evaluate(self, points, classification=False):
        """Classify new samples then predict with the matching local model (mixture.py:305-353)."""
        points = self.scaler.transform(np.atleast_2d(np.asarray(points, dtype=np.float64)))
        classif = self.classifier.predict(points)
        result = np.empty((len(points), self.fsizes))
        sigma = np.empty((len(points), self.fsizes))
        for k in np.unique(classif):
            sel = np.flatnonzero(classif == k)
            result[sel], sigma[sel] = self.local_models[k](points[sel])
        return (result, sigma, classif) if classification else (result, sigma)
