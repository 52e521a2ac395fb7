"""Stand-in for ftfy.fix_text (simple_tokenizer.py): the ImageNet prompts are plain ASCII, nothing to fix."""


def fix_text(text, **kwargs):
    return text
