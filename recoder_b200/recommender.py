"""Recommender adapters with the reference's interface (recoder/recommender.py:8-24, 104-118).  The Annoy-based
`SimilarityRecommender` (recommender.py:27-101) is a serving-side component outside the training path
 and is not rebuilt.

Attribution: the public interface of this module (class / method names, argument lists and their documentation, log
messages, checkpoint keys) mirrors amoussawi/recoder (MIT License, Copyright (c) 2018 Abdallah Moussawi) so that it is
a drop-in for that library; see LICENSE.  The implementation underneath is original.
"""


class Recommender(object):
  """Base class for recommenders: ``recommend(users_hist) -> list of item-id lists``."""

  def recommend(self, users_hist):
    raise NotImplementedError


class InferenceRecommender(Recommender):
  """
  Recommends items based on the predictions by a :class:`recoder_b200.model.Recoder` model.

  Args:
    model (Recoder): model used to predict recommendations
    num_recommendations (int): number of recommendations to generate for each user.
  """

  def __init__(self, model, num_recommendations):
    self.model = model
    self.num_recommendations = num_recommendations

  def recommend(self, users_hist):
    return self.model.recommend(users_hist, self.num_recommendations)
