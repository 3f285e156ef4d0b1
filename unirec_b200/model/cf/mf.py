"""Matrix factorisation (drop-in for unirec/model/cf/mf.py:6-8): user = U[user_id]; everything else is BaseRecommender."""
from unirec_b200.model.base.recommender import BaseRecommender


class MF(BaseRecommender):
    _tower_kind = 'mf'
