"""See ../__init__.py: names exist so that ``from nltk.translate.bleu_score import SmoothingFunction`` succeeds."""


def _missing(*_a, **_k):
    raise ImportError("the BLEU metric needs the real `nltk` package; this image only carries an import-time stand-in")


class SmoothingFunction:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        if name.startswith("method"):
            return _missing
        raise AttributeError(name)


sentence_bleu = corpus_bleu = _missing
