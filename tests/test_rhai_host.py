"""Host-side tests of the script interpreter (paintfe_b200/rhai_host.py), modelled on the reference's
tests/scripting.rs: canvas / pixel / utility / selection API and the language around it.

No GPU: the device entry points are replaced by an engine-shaped stand-in over the oracle, so only the host logic
(parser, evaluator, closure evaluation over whole images vs per pixel, image / mask bookkeeping) is under test.
The same scripts run against the CUDA path in test_gpu_parity.py::test_script_host_on_device.
"""
import numpy as np
import pytest

import fixtures as fx
from paintfe_b200.rhai_host import Interpreter, Parser, ScriptError, run_script


class OracleEngine:
    """The few Engine methods these scripts reach, answered by the oracle."""

    def __init__(self, pfo):
        self.o = pfo

    def canvas_border(self, img, width, color, mask=None):
        return self.o.canvas_border(img, width, color, mask=mask)

    def adjust(self, img, op, params=(), luts=None):
        return self.o.adjust(img, op, params, luts=luts)

    def gaussian_blur(self, img, sigma, mask=None, exact=False):
        return self.o.gaussian_blur(img, sigma, mask=mask)

    def pixelate(self, img, size, mask=None):
        return self.o.pixelate(img, size, mask=mask)

    def orient(self, img, op):
        return self.o.orient(img, op)

    def resize(self, img, w, h, filt):
        return self.o.resize(img, w, h, filt)

    def resize_canvas(self, img, w, h, anchor, fill):
        return self.o.resize_canvas(img, w, h, anchor, fill)

    def vignette(self, img, amount, softness, mask=None):
        return self.o.vignette(img, amount, softness, mask=mask)


@pytest.fixture()
def run(oracle):
    eng = OracleEngine(oracle)

    def go(source, img=None, mask=None):
        return run_script(eng, source, fx.gradient(64, 64) if img is None else img, mask)

    return go


# ---- Canvas / pixel API (tests/scripting.rs:31-84)
def test_width_height(run):
    _, console = run("let w = width(); let h = height(); print_line(`${w}x${h}`);")
    assert console[-1] == "64x64"


def test_set_pixel(run):
    src = fx.gradient(64, 64)
    keep = src.copy()
    out, _ = run("set_pixel(0, 0, 255, 0, 0, 255); set_pixel(1, 0, 0, 255, 0, 128); set_pixel(-1, 0, 1, 1, 1, 1); set_pixel(64, 0, 1, 1, 1, 1);", src)
    assert out[0, 0].tolist() == [255, 0, 0, 255] and out[0, 1].tolist() == [0, 255, 0, 128]
    assert (src == keep).all(), "the caller's buffer is not written"
    assert (out[1:] == keep[1:]).all()


def test_get_pixel_roundtrip(run):
    _, console = run("""
        set_pixel(5, 5, 10, 20, 30, 40);
        let p = get_pixel(5, 5);
        print_line(`${p[0]},${p[1]},${p[2]},${p[3]}`);
        let q = get_pixel(-3, 900);
        print_line(`${q}`);
        set_g(5, 5, 999); print_line(`${get_g(5, 5)} ${get_a(5, 5)} ${get_r(64, 0)}`);
    """)
    assert console == ["10,20,30,40", "[0, 0, 0, 0]", "255 40 0"]


# ---- Bulk iteration against the reference's goldens (tests/scripting.rs:88-113)
def test_for_each_pixel_invert_golden(run):
    out, _ = run("for_each_pixel(|x, y, r, g, b, a| { [255 - r, 255 - g, 255 - b, a] });")
    assert (out == fx.golden("scripting", "for_each_pixel_invert")).all()


def test_map_channels_invert_golden(run):
    out, _ = run("map_channels(|r, g, b, a| { [255 - r, 255 - g, 255 - b, a] });")
    assert (out == fx.golden("scripting", "map_channels_invert")).all()


def test_script_invert_matches_native(run):
    a, _ = run("apply_invert();")
    assert (a == fx.golden("scripting", "apply_invert")).all()
    b, _ = run("map_channels(|r, g, b, a| [255 - r, 255 - g, 255 - b, a]);")
    assert (a == b).all()


def test_for_region_and_unit_results(run):
    src = fx.gradient(64, 64)
    out, _ = run("for_region(60, -2, 10, 6, |x, y, r, g, b, a| { if x == 62 { return; } [0, 0, y, a] });", src)
    exp = src.copy()
    for y in range(0, 4):
        for x in (60, 61, 63):
            exp[y, x] = [0, 0, y, 255]
    assert (out == exp).all()


def test_closure_float_results_keep_the_channel(run):
    """`arr[i].as_int().unwrap_or(old)` (scripting.rs:466): a float result leaves the channel as it was."""
    src = fx.gradient(64, 64)
    out, _ = run("map_channels(|r, g, b, a| [r / 2.0, to_int(floor(g / 2.0)), 300, -5]);", src)
    assert (out[..., 0] == src[..., 0]).all() and (out[..., 1] == src[..., 1] // 2).all()
    assert (out[..., 2] == 255).all() and (out[..., 3] == 0).all()


def _per_pixel_only(body):
    """The same closure body forced through the per-pixel evaluator (rand_* depends on the visiting order)."""
    return "{ let _order = rand_int(0, 2); " + body[1:]


@pytest.mark.parametrize("body", [
    "{ let v = (r * 3 + g) / 4; if v > 100 { v = 100; } else if v < 10 { v -= 5; } [v, g % 7, b / 3, a] }",
    "{ if (x + y) % 2 == 0 && is_selected(x, y) { [b, r, g, a] } else { [r, g, b, 255 - a] } }",
    "{ let d = distance(x, y, 32, 32); let k = clamp(to_int(d * 4.0), 0, 255); [k, max(r, g), min(r, b), abs(r - g)] }",
    "{ let t = [r, g]; if r > g { t[0] = g; t[1] = r; } [t[0], t[1], to_int(round(lerp(r, b, 0.5))), a] }",
    "{ let q = -r / 3 + (-g) % 5 + 2 ** 3; [q + 100, r & 15 | 64, (g ^ b) >> 1, a] }",
])
def test_whole_image_and_per_pixel_evaluation_agree(run, body):
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, (24, 40, 4), dtype=np.uint8)
    script = "select_ellipse(20.0, 12.0, 15.0, 9.0); for_each_pixel(|x, y, r, g, b, a| %s);"
    fast, _ = run(script % body, src)
    slow, _ = run(script % _per_pixel_only(body), src)
    assert (fast == slow).all()
    assert not (fast == src).all()


def test_select_rect_then_closure(run):
    out, _ = run("""
        select_rect(0, 0, 32, 64);
        for_each_pixel(|x, y, r, g, b, a| {
            if is_selected(x, y) { [255 - r, 255 - g, 255 - b, a] } else { [r, g, b, a] }
        });
    """)
    assert out[32, 5, 0] > 200 and out[32, 50, 0] > 100
    src = fx.gradient(64, 64)
    assert (out[:, :32, :3] == 255 - src[:, :32, :3]).all() and (out[:, 32:] == src[:, 32:]).all()


# ---- Utility API (tests/scripting.rs:188-213)
def test_print(run):
    _, console = run('print_line("hello world"); print("second line"); print(42); print(1.0); print(0.25); print(true); print([1, "a", 2.5]);')
    assert console == ["hello world", "second line", "42", "1.0", "0.25", "true", '[1, "a", 2.5]']


def test_math_functions(run):
    _, console = run("""
        let v = clamp(300, 0, 255); print_line(`${v}`);
        print(7 / 2); print(-7 / 2); print(-7 % 3); print(7.0 / 2); print(2 ** 10); print(pow(2.0, 0.5));
        print(floor(2.7)); print(ceil(2.1)); print(round(2.5)); print(round(-2.5)); print(sqrt(16.0));
        print(min(3, 4)); print(max(1.5, 2.5)); print(abs(-3)); print(lerp(0.0, 10.0, 0.25)); print(distance(0.0, 0.0, 3.0, 4.0));
        print(rgb_to_hsl(255, 0, 0)); print(hsl_to_rgb(120.0, 100.0, 50.0)); print(hsl_to_rgb(0.0, 0.0, 50.0));
        print(to_int(3.9)); print(to_float(3)); print((2.9).to_int()); print(PI() > 3.14 && PI() < 3.15);
    """)
    assert console == ["255", "3", "-3", "-1", "3.5", "1024", repr(2 ** 0.5), "2.0", "3.0", "3.0", "-3.0", "4.0",
                       "3", "2.5", "3", "2.5", "5.0", "[0.0, 100.0, 50.0]", "[0, 255, 0]", "[128, 128, 128]",
                       "3", "3.0", "2", "true"]


def test_rand_is_the_reference_xorshift(oracle):
    """scripting.rs:1216-1256: xorshift64 (13, 7, 17) on the context's state."""
    it = Interpreter(OracleEngine(oracle), fx.gradient(4, 4), seed=88172645463325252)
    it.run("print(rand_int(0, 1000)); print(rand_int(5, 5)); let f = rand_float(); print(f >= 0.0 && f <= 1.0); print(rand_float(2.0, 1.0));")
    s = 88172645463325252
    s ^= (s << 13) & (2 ** 64 - 1); s ^= s >> 7; s ^= (s << 17) & (2 ** 64 - 1)
    assert it.console == [str(s % 1000), "5", "true", "2.0"]


# ---- Language
def test_control_flow_functions_and_arrays(run):
    _, console = run("""
        fn lum(r, g, b) { (r * 30 + g * 59 + b * 11) / 100 }
        fn fact(n) { if n <= 1 { return 1; } n * fact(n - 1) }
        let n = 0; let seen = [];
        while n < 5 { n += 1; if n == 2 { continue; } if n == 5 { break; } seen.push(n); }
        print(seen);
        let total = 0;
        for i in 0..4 { total += i; } for i in 0..=4 { total += i; } for i in range(10, 0, -5) { total += i; } for v in [100, 200] { total += v; }
        print(total);
        let k = 0; loop { k += 1; if k >= 3 { break; } } print(k);
        print(lum(255, 255, 255)); print(fact(10)); print(seen.len()); print(seen[-1]); print(seen.contains(3));
        let label = if total > 10 { "big" } else { "small" }; print(label + "!" + 1);
        const Z = 0x10; print(Z); /* block
        comment */ print(1_000);
        let f = |a, b| a * b; print(f.call(6, 7));
        let s = "tab\\there"; print(s.len());
    """)
    assert console == ["[1, 3, 4]", "331", "3", "255", "3628800", "3", "4", "true", "big!1", "16", "1000", "42", "8"]


@pytest.mark.parametrize("source", ["let x = ;", "let x = 1 / 0;", "foo(1);", "let a = [1]; a[3];", "print(y);", "if 1 { }",
                                    "for_each_pixel(3);", "let x = 1 +;", "while true { ", "apply_blur(|x| x);", "1 + true;",
                                    "`${`", "let x = 5 % 0;", "break;", "fn f(n) { f(n) } f(1);"])
def test_errors_are_script_errors(run, source):
    """tests/scripting.rs:217-233: syntax and runtime errors carry a message."""
    with pytest.raises(ScriptError) as err:
        run(source)
    assert err.value.message


def test_error_carries_the_line(run):
    with pytest.raises(ScriptError) as err:
        run("let a = 1;\nlet b = 2;\nlet c = ;")
    assert err.value.line == 3 and "Line 3" in str(err.value)


def test_operation_limit(oracle):
    it = Interpreter(OracleEngine(oracle), fx.gradient(4, 4))
    it.MAX_OPERATIONS = 10_000
    with pytest.raises(ScriptError, match="too many operations"):
        it.run("let i = 0; loop { i += 1; }")


# ---- Selection API (tests/scripting.rs:266-420)
def test_selection_api(run):
    out, _ = run("select_rect(10, 10, 30, 30); fill_selected(255, 0, 0, 255);")
    assert out[20, 20].tolist() == [255, 0, 0, 255] and out[5, 5, 0] != 255
    out, _ = run("select_ellipse(32.0, 32.0, 15.0, 15.0); fill_selected(255, 0, 255, 255);")
    assert out[32, 32].tolist()[:3] == [255, 0, 255] and out[0, 0].tolist()[:2] == [0, 255]
    out, _ = run("select_rect(0, 0, 10, 10); clear_selection(); fill_selected(0, 0, 255, 255);")
    assert (out[..., 2] == 255).all()
    _, console = run('print_line("before: " + has_selection()); select_rect(0, 0, 10, 10); print_line("after: " + has_selection());'
                     ' clear_selection(); print_line("cleared: " + has_selection());')
    assert console == ["before: false", "after: true", "cleared: false"]
    out, _ = run("select_rect(10, 10, 54, 54); invert_selection(); fill_selected(255, 0, 255, 255);")
    assert out[0, 0, 0] == 255 and out[0, 0, 2] == 255 and (out[32, 32, 0], out[32, 32, 2]) != (255, 255)
    out, _ = run("select_rect(20, 20, 44, 44); delete_selected();")
    assert out[32, 32, 3] == 0 and out[5, 5, 3] > 0


def test_effects_mix_with_pixel_access(run, oracle):
    """set_pixel before an effect is seen by it; get_pixel after it sees the result; a resize drops the selection."""
    out, console = run("""
        set_pixel(0, 0, 0, 0, 0, 255);
        apply_invert();
        let p = get_pixel(0, 0); print(p);
        select_rect(0, 0, 8, 8);
        resize_image(32, 32, "nearest");
        print(`${width()}x${height()} ${has_selection()}`);
        if width() < 64 { flip_horizontal(); }
    """)
    assert console == ["[255, 255, 255, 255]", "32x32 false"]
    src = fx.gradient(64, 64)
    src[0, 0] = [0, 0, 0, 255]
    src[..., :3] = 255 - src[..., :3]
    assert (out == oracle.orient(oracle.resize(src, 32, 32, 0), 0)).all()


def test_parser_keeps_statement_and_expression_if_apart():
    prog = Parser("if a { 1 } [2]; let v = if a { 1 } else { 2 };").program()
    assert [st[0] for st in prog] == ["expr", "expr", "let"] and prog[1][1][0] == "arr"


def test_canvas_ops_are_logged_and_replayed(oracle):
    """scripting.rs:687-813 log, :1640 replay: layer-only flips log nothing, an unchanged-size resize logs nothing."""
    from paintfe_b200.script import apply_canvas_ops

    eng = OracleEngine(oracle)
    rng = np.random.default_rng(9)
    active, other = fx.random_rgba(rng, 48, 32), fx.random_rgba(rng, 48, 32)
    it = Interpreter(eng, active)
    out = it.run('flip_horizontal(); rotate_180(); rotate_canvas_90cw(); resize_image(32, 48, "nn"); resize_image(16, 24, "lanczos");'
                 ' resize_canvas(20, 30, "center"); flip_canvas_vertical();')
    assert [op[0] for op in it.canvas_ops] == ["rotate_canvas_90cw", "resize_image", "resize_canvas", "flip_canvas_vertical"]
    flats = apply_canvas_ops(eng, [other, out, None], 1, it.canvas_ops)
    exp = oracle.orient(oracle.resize_canvas(oracle.resize(oracle.orient(other, oracle.ROT90CW), 16, 24, 3), 20, 30, (1, 1), (0, 0, 0, 0)), oracle.FLIP_V)
    assert flats[1] is out and flats[2] is None and (flats[0] == exp).all() and out.shape == exp.shape


# ---- Property: evaluating a closure once over whole-image arrays == evaluating it per pixel
def _random_body(rng, depth=0):
    """A random closure body over r, g, b, a, x, y: integer / float arithmetic, comparisons, data-dependent ifs."""
    ints = ["r", "g", "b", "a", "x", "y", "7", "255", "3", "(r + g)", "(b - a)", "(x * 2 + y)"]

    def int_expr(d):
        if d > 2 or rng.random() < 0.3:
            return ints[rng.integers(len(ints))]
        op = ["+", "-", "*", "/", "%", "&", "|", "^"][rng.integers(8)]
        rhs = int_expr(d + 1)
        if op in "/%":
            rhs = f"({rhs} | 1)"  # never zero
        kind = rng.integers(6)
        if kind == 0:
            return f"max({int_expr(d + 1)}, {rhs})"
        if kind == 1:
            return f"clamp({int_expr(d + 1)}, 0, 255)"
        if kind == 2:
            return f"to_int(floor({int_expr(d + 1)} / 2.5))"
        if kind == 3:
            return f"abs({int_expr(d + 1)} - {rhs})"
        return f"({int_expr(d + 1)} {op} {rhs})"

    def cond(d):
        c = f"{int_expr(d)} {['<', '<=', '>', '>=', '==', '!='][rng.integers(6)]} {int_expr(d)}"
        if rng.random() < 0.3:
            c = f"({c}) {['&&', '||'][rng.integers(2)]} is_selected(x, y)"
        return c

    lines = [f"let u = {int_expr(0)};", f"let v = {int_expr(0)};"]
    for _ in range(rng.integers(1, 4)):
        if rng.random() < 0.5:
            lines.append(f"if {cond(0)} {{ u = {int_expr(1)}; }} else if {cond(1)} {{ v += {int_expr(1)}; }} else {{ u -= 1; v = u; }}")
        else:
            lines.append(f"let w = if {cond(0)} {{ {int_expr(1)} }} else {{ {int_expr(1)} }}; u = u + w % 17;")
    lines.append(f"[u, v, if {cond(0)} {{ b }} else {{ 255 - b }}, {['a', '255', 'a / 2.0'][rng.integers(3)]}]")
    return "{ " + " ".join(lines) + " }"


@pytest.mark.parametrize("seed", range(25))
def test_random_closures_whole_image_vs_per_pixel(oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    body = _random_body(rng)
    src = rng.integers(0, 256, (9, 13, 4), dtype=np.uint8)
    script = "select_rect(2, 1, 11, 7); for_each_pixel(|x, y, r, g, b, a| %s);"
    fast, slow = Interpreter(OracleEngine(oracle), src), Interpreter(OracleEngine(oracle), src)
    a, b = fast.run(script % body), slow.run(script % _per_pixel_only(body))
    assert fast.bulk_evaluations == {"whole_image": 1, "per_pixel": 0}, body
    assert slow.bulk_evaluations == {"whole_image": 0, "per_pixel": 1}
    assert (a == b).all(), body


def test_reference_readme_script(run, oracle):
    """The example of the reference's README (Scripting section), statement for statement."""
    out, _ = run("""
        apply_desaturate();
        apply_brightness_contrast(10.0, 40.0);
        apply_vignette(0.5, 0.3);

        map_channels(|r, g, b, a| {
            [clamp(r + 15, 0, 255), g, clamp(b - 8, 0, 255), a]
        });
    """)
    exp = oracle.adjust(fx.gradient(64, 64), 33)                       # PFE_ADJ_S_DESATURATE
    exp = oracle.adjust(exp, 36, (10.0, 40.0))                         # PFE_ADJ_S_BRIGHTNESS_CONTRAST
    exp = oracle.vignette(exp, 0.5, 0.3).astype(np.int64)
    exp[..., 0] = np.clip(exp[..., 0] + 15, 0, 255)
    exp[..., 2] = np.clip(exp[..., 2] - 8, 0, 255)
    assert (out == exp.astype(np.uint8)).all()


# ---------------------------------------------------------------------------------------------------------------
# closures that write outside their own frame: Rhai shares captured variables and runs the closure once per pixel, so
# counters, running sums, histograms and captured arrays must see every pixel - the whole-image evaluation has to
# decline them (and leave no trace of its attempt behind)
# ---------------------------------------------------------------------------------------------------------------
def _host_only(src, img):
    from paintfe_b200.rhai_host import Interpreter

    it = Interpreter(None, img.copy())
    out = it.run(src)
    return it, out


def test_impure_closures_are_evaluated_per_pixel():
    img = (np.arange(20 * 4, dtype=np.uint8).reshape(4, 5, 4) * 3).astype(np.uint8)
    it, _ = _host_only("let count = 0; for_each_pixel(|x,y,r,g,b,a| { count += 1; [r,g,b,a] }); print(count);", img)
    assert it.console == ["20"] and it.bulk_evaluations == {"whole_image": 0, "per_pixel": 1}
    it, _ = _host_only("let t = 0; for_each_pixel(|x,y,r,g,b,a| { t = t + r; [r,g,b,a] }); print(t);", img)
    assert it.console == [str(int(img[..., 0].astype(int).sum()))]
    it, _ = _host_only("let hist = [0, 0]; map_channels(|r,g,b,a| { if r > 100 { hist[1] += 1 } else { hist[0] += 1 }; [r,g,b,a] }); print(hist);", img)
    hi = int((img[..., 0] > 100).sum())
    assert it.console == [f"[{20 - hi}, {hi}]"]
    # mutation BEFORE the construct that forces the per-pixel path: nothing of the abandoned attempt may survive
    it, _ = _host_only("let acc = []; let count = 0; for_each_pixel(|x,y,r,g,b,a| { count += 1; let v = rand_int(0, 10); acc.push(v); [r,g,b,a] });"
                       " print(count); print(acc.len());", img)
    assert it.console == ["20", "20"]
    # the random sequence starts where it would have without the attempt
    a, _ = _host_only("let s = 0; for_each_pixel(|x,y,r,g,b,a| { s += rand_int(0, 1000); [r,g,b,a] }); print(s);", img)
    b, _ = _host_only("let s = 0; for x in 0..20 { s += rand_int(0, 1000); } print(s);", img)
    assert a.console == b.console
    # a pure closure still takes the whole-image path, and gives the per-pixel result
    it, out = _host_only("for_each_pixel(|x,y,r,g,b,a| { let k = 255 - r; [k, g, b, a] });", img)
    assert it.bulk_evaluations == {"whole_image": 1, "per_pixel": 0} and np.array_equal(out[..., 0], 255 - img[..., 0])


def test_integer_arithmetic_is_checked_like_rhai():
    from paintfe_b200.rhai_host import ScriptError

    img = np.zeros((2, 2, 4), np.uint8)
    for src in ("let x = 9223372036854775807; x += 1;", "let x = 3037000500; let y = x * x;", "let z = 2 ** 70;",
                "let m = -9223372036854775807 - 2;"):
        with pytest.raises(ScriptError):
            _host_only(src, img)
    it, _ = _host_only("print(2 ** 62); print(9223372036854775807 - 1);", img)
    assert it.console == ["4611686018427387904", "9223372036854775806"]
    # products that would wrap in int64 arrays leave the whole-image path instead of wrapping
    it, out = _host_only("for_each_pixel(|x,y,r,g,b,a| { let k = ((r + 7) * 3000000000) / 3000000000; [k, g, b, a] });", img)
    assert np.all(out[..., 0] == 7)
