"""Unit tests of oracle/minijs, the interpreter that runs the unmodified main.js for the golden
frames: the JS semantics main.js leans on must hold in it."""
import math

from oracle.minijs.interp import Interpreter, JSArray, Uint8ClampedArray, to_string


def run(src):
    it = Interpreter()
    it.run(src)
    it.drain_timers()
    return it.globals.vars


def test_numbers_closures_and_hoisting():
    g = run("""
      'use strict';
      const a = 1/3 + 2*0.1;            // plain IEEE doubles
      let s = (0.1 + 0.2) * 10;
      function outer(k) { let n = 0; function inc() { n += k; return n; } return inc; }
      const f = outer(2.5); f(); const r = f();
      const hoisted = later(4);
      function later(x) { return x * x; }
      const t = (5000.75 & 1) ^ (2500.2 & 1);   // & goes through ToInt32
      const big = (4294967301 & 7);
      let i = 0; const arr = []; arr[i++] = 7; arr[i++] = 8; arr[5] = 1;
      const cond = (a > 0.5) ? 'hi' : 'lo';
    """)
    assert g["a"] == 1 / 3 + 2 * 0.1 and g["s"] == (0.1 + 0.2) * 10
    assert g["r"] == 5.0 and g["hoisted"] == 16.0
    assert g["t"] == 0.0 and g["big"] == 5.0
    assert g["i"] == 2.0 and len(g["arr"].items) == 6 and g["arr"].items[1] == 8.0
    assert g["cond"] == "hi"


def test_stable_sort_and_array_methods():
    g = run("""
      const xs = [{k:1,v:'a'},{k:0,v:'b'},{k:1,v:'c'},{k:0,v:'d'}];
      xs.sort(function (p, q) { return p.k - q.k; });
      let out = '';
      xs.forEach(function (o, i) { out += o.v + i; });
      const j = [1, 'x', 2.5].join('-');
    """)
    assert g["out"] == "b0d1a2c3"            # ties keep insertion order (main.js:159-163 relies on it)
    assert g["j"] == "1-x-2.5"


def test_clamped_array_store():
    a = Uint8ClampedArray(8)
    for i, v in enumerate([math.nan, -5, 0.5, 1.5, 2.5, 254.5, 255.7, 1e9]):
        a.js_set(float(i), v)
    assert list(a.buf) == [0, 0, 0, 2, 2, 254, 255, 255]


def test_timers_this_and_rest_params():
    g = run("""
      let order = '';
      setTimeout(function (x, y) { order += 'A' + x + y; }, 0, 1, 2);
      setTimeout(function () { order += 'B'; setTimeout(function () { order += 'D'; }, 500); }, 0);
      setTimeout(function () { order += 'C'; }, 0);
      const o = { n: 3, get: function (...rest) { return this.n + rest.length; } };
      const viaMethod = o.get(1, 2, 3);
      function f() { return typeof this; }
      const bare = f();
      const inf = 1 / 0; const nan = 0 / 0; const cmp = (nan < 1) || (nan >= 1);
      const mx = Math.max(-1, Math.min(1, 2)); const mnan = Math.min(1, nan);
    """)
    assert g["order"] == "A12BCD"            # FIFO drain, nested timers appended
    assert g["viaMethod"] == 6.0 and g["bare"] == "undefined"
    assert g["inf"] == math.inf and g["nan"] != g["nan"] and g["cmp"] is False
    assert g["mx"] == 1.0 and g["mnan"] != g["mnan"]
    assert to_string(12.0) == "12" and to_string(0.5) == "0.5"
