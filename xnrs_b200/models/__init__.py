"""Mirror of the reference's ``xnrs.models`` package surface (xnrs/models/__init__.py, make_model.py)."""
from .components import (AdditiveAttention, BilinScoring, DotScoring, FCScoring, MaskedMean, MultiHeadAttention, ParentRec,
                         PersonalizedAttention, TextEncoder, UserEncoder)
from .zoo import (LSTUR, NAML, NPA, NRMS, NRMS_LF, BaseRec, LSTURNewsEncoder, LSTURUserEncoder, MeanRec, ParamFreeRec, SmallNAML,
                  StandardRec, make_model)

__all__ = ['AdditiveAttention', 'BilinScoring', 'DotScoring', 'FCScoring', 'MaskedMean', 'MultiHeadAttention', 'ParentRec',
           'PersonalizedAttention', 'TextEncoder', 'UserEncoder', 'LSTUR', 'NAML', 'NPA', 'NRMS', 'NRMS_LF', 'BaseRec', 'MeanRec',
           'ParamFreeRec', 'SmallNAML', 'LSTURNewsEncoder', 'LSTURUserEncoder', 'StandardRec', 'make_model']
