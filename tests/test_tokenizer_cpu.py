"""Native BPE tokenizer (csrc/tokenizer.cu) against the REAL reference tokenizer
(lib/dataset/languages/simple_tokenizer.py, imported from /root/reference or the staged baseline/_ref with ftfy stubbed to the
identity - it is not installed here, and our wrapper then skips it too): integer work, so ids must be IDENTICAL.  The merges
file is the reference's own data; the tests skip where neither copy of the reference is present."""
import importlib.util
import os
import random
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = ["/root/reference/lib/dataset/languages", os.path.join(ROOT, "baseline", "_ref", "lib", "dataset", "languages")]
LANG_DIR = next((d for d in CANDIDATES if os.path.exists(os.path.join(d, "bpe_simple_vocab_16e6.txt.gz"))), None)
pytestmark = pytest.mark.skipif(LANG_DIR is None, reason="needs the reference's bpe_simple_vocab_16e6.txt.gz")


@pytest.fixture(scope="module")
def pair():
    if "ftfy" not in sys.modules:
        stub = types.ModuleType("ftfy")
        stub.fix_text = lambda t: t
        sys.modules["ftfy"] = stub
    spec = importlib.util.spec_from_file_location("ref_simple_tokenizer", os.path.join(LANG_DIR, "simple_tokenizer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ref = mod.SimpleTokenizer(os.path.join(LANG_DIR, "bpe_simple_vocab_16e6.txt.gz"))
    from msclip_b200.tokenizer import SimpleTokenizer
    ours = SimpleTokenizer(os.path.join(LANG_DIR, "bpe_simple_vocab_16e6.txt.gz"))
    return ref, ours


PROMPTS = [
    "a photo of a cat.", "A Photo Of The LARGE tench, a type of fish!", "itap of a    golden   retriever\t\n", "", "   ",
    "don't can't we'll they've I'm you'd it's 'tis rock'n'roll", "x2 3d 1080p 4k8k 12345 3.14159 1,000,000", "hello---world!!! ??? ... (test) [ok] {z}",
    "&lt;tag&gt; &amp;amp; café naïve über straße", "你好世界 こんにちは 안녕", "emoji \U0001F600\U0001F680 ok ❤️",
    "İstanbul İ ΣΟΦΟΣ σοφόΣ. ΟΣ Σ", "it'ſ a'ſk 'ſ", "<|startoftext|> hi <|endoftext|> <|STARTOFTEXT|>",
    "tab\tnew\nline\r\nnbsp thin zero​width \x1c\x1d", "Ⅷ ½ ٣٤ ① numbers", "a" * 300, " ".join(["word"] * 120),
    "não é combining", "K kelvin Ω ohm ﬁ ligature ẞ", "'s 't 're 've 'm 'll 'd ' '' 'x",
]


def test_encode_matches_the_reference_on_hand_picked_prompts(pair):
    ref, ours = pair
    for text in PROMPTS:
        assert ours.encode(text) == ref.encode(text), repr(text)


def test_tokenize_matches_the_reference_incl_truncation_and_padding(pair):
    ref, ours = pair
    a, b = ours.tokenize(PROMPTS), ref.tokenize(PROMPTS)
    assert a.dtype == torch.long and a.shape == b.shape == (len(PROMPTS), 77)
    assert torch.equal(a, b)
    assert torch.equal(ours("a photo of a dog", 16), ref("a photo of a dog", 16))
    assert (ours.get_sot_token(), ours.get_eot_token(), ours.get_vocab_size()) == (ref.get_sot_token(), ref.get_eot_token(), ref.get_vocab_size())


def test_random_unicode_soup_matches_the_reference(pair):
    """Random strings over a pool of awkward code points (case pairs, marks, digits of several scripts, spaces of all kinds,
    punctuation, astral planes): every id must agree."""
    ref, ours = pair
    pool = [chr(c) for c in list(range(0x20, 0x7f)) + list(range(0xa0, 0x180)) + list(range(0x370, 0x400)) + list(range(0x400, 0x460)) +
            [0x130, 0x131, 0x17f, 0x1c5, 0x1c8, 0x2bc, 0x300, 0x301, 0x307, 0x345, 0x3a3, 0x3c2, 0x3c3, 0x660, 0x966, 0x1680, 0x2000, 0x2009,
             0x200b, 0x2028, 0x202f, 0x205f, 0x2126, 0x212a, 0x2160, 0x2460, 0x3000, 0x4e00, 0x4e8c, 0xac00, 0xfb01, 0xff21, 0xff41, 0x1d400,
             0x1f600, 0x10400, 0x10428, 0x1e900, 0x1e922, 0x9, 0xa, 0xd, 0x1c, 0x1f, 0x85]]
    rng = random.Random(1234)
    texts = ["".join(rng.choice(pool) for _ in range(rng.randint(0, 60))) for _ in range(1500)]
    got = ours.tokenize(texts, 77)
    want = ref.tokenize(texts, 77)
    bad = [i for i in range(len(texts)) if not torch.equal(got[i], want[i])]
    assert not bad, [(repr(texts[i]), ours.encode(texts[i]), ref.encode(texts[i])) for i in bad[:3]]


def test_zero_shot_prompt_set_is_identical_and_faster(pair):
    """The tool's workload (tools/zero_shot.py:121-132): class names x 80 templates in one call."""
    import time
    ref, ours = pair
    names = ["tench", "goldfish", "great white shark", "tiger shark", "hammerhead", "electric ray", "stingray", "cock", "hen", "ostrich"] * 10
    templates = ["a bad photo of a {}.", "a photo of many {}.", "a sculpture of a {}.", "a photo of the hard to see {}.",
                 "a low resolution photo of the {}.", "a rendering of a {}.", "graffiti of a {}.", "a bad photo of the {}."] * 10
    texts = [t.format(n) for n in names for t in templates]
    t0 = time.perf_counter()
    want = ref.tokenize(texts)
    t1 = time.perf_counter()
    got = ours.tokenize(texts)
    t2 = time.perf_counter()
    assert torch.equal(got, want)
    print(f"{len(texts)} prompts: reference {t1 - t0:.3f} s, native {t2 - t1:.3f} s")


def test_dropin_swaps_the_tokenizer_class_of_the_reference_package(pair):
    """dropin.install(tokenizer=True): `from dataset.languages import SimpleTokenizer` (tools/zero_shot.py:34) then yields the
    native tokenizer with the reference's default merges file; methods outside the zero-shot path fall through."""
    import subprocess
    lib_dir = os.path.dirname(os.path.dirname(LANG_DIR))
    code = f"""
import sys, types
stub = types.ModuleType('ftfy'); stub.fix_text = lambda t: t; sys.modules['ftfy'] = stub
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {lib_dir!r})
import msclip_b200.dropin as d
d.install(tokenizer=True)
from dataset.languages import SimpleTokenizer
import msclip_b200.tokenizer as T
t = SimpleTokenizer()
assert isinstance(t, T.SimpleTokenizer), type(t)
ids = t('a photo of a dog')[0, :7].tolist()
assert ids[0] == t.get_sot_token() and t.get_eot_token() in ids, ids
assert t.decode(t.encode('hello world')).strip() == 'hello world'
print('OK', ids)
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
