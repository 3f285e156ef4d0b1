from unirec_b200.model.base.recommender import BaseRecommender


class SeqRecBase(BaseRecommender):
    """Common base of the sequence models (reference: unirec/model/sequential/seqrec_base.py:10-27)."""

    def add_annotation(self):
        super().add_annotation()
        self.annotations.append('SeqRecBase')
