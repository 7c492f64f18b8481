"""Mirror of the reference's ``xnrs.models`` package surface (xnrs/models/__init__.py, make_model.py)."""
from .components import (AdditiveAttention, DotScoring, MaskedMean, MultiHeadAttention, ParentRec,
                         PersonalizedAttention, TextEncoder, UserEncoder)
from .zoo import LSTUR, NAML, NPA, NRMS, LSTURNewsEncoder, LSTURUserEncoder, StandardRec, make_model

__all__ = ['AdditiveAttention', 'DotScoring', 'MaskedMean', 'MultiHeadAttention', 'ParentRec',
           'PersonalizedAttention', 'TextEncoder', 'UserEncoder', 'LSTUR', 'NAML', 'NPA', 'NRMS',
           'LSTURNewsEncoder', 'LSTURUserEncoder', 'StandardRec', 'make_model']
