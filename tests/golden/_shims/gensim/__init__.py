"""Import stub: fork models import gensim eagerly (general_recommender/jointsr.py:7-8)."""
