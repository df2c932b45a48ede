"""gravity/survey.py mirror."""
from .._survey import BaseSurvey


class Survey(BaseSurvey):
    pass
