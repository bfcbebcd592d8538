#!/usr/bin/env node
// run_main_js.js — Node `vm` harness that executes the UNMODIFIED reference main.js headless and dumps
// its framebuffer (SURVEY.md §8c, plan item 3).  TEST INFRASTRUCTURE; NOT EXECUTED in the build image
// (no Node there) — oracle/minijs + gen_golden.py play this role until a box with Node shows up.
//
//   node oracle/run_main_js.js <reference-dir> <width> <height> <textures.json> <out.rgba> [--random 0.5]
//
// textures.json: {"earth.png": {"width":256,"height":128,"rgba":"<base64 of W*H*4 bytes>"}, "mars.png": {...}}
// (the decoded bytes getImageData would return; write it with tools/make_assets.py --json).
// Output: raw RGBA8 (W*H*4 bytes) + a JSON line on stdout with the overlay text and timings.
'use strict';
const fs = require('fs');
const path = require('path');
const vm = require('vm');

const [refDir, W, H, texJson, outPath] = process.argv.slice(2);
const randomIdx = process.argv.indexOf('--random');
const randomValue = randomIdx > 0 ? Number(process.argv[randomIdx + 1]) : 0.5;   // policy P0: sky black
if (!outPath) { console.error('usage: run_main_js.js <reference-dir> <W> <H> <textures.json> <out.rgba>'); process.exit(2); }
const width = Number(W), height = Number(H);
const textures = JSON.parse(fs.readFileSync(texJson, 'utf8'));

// ---- synchronous FIFO for setTimeout: removes timer latency, makes the run deterministic ----------------
const queue = [];
const captured = { puts: 0, text: null, imagedata: null };

function makeContext2d() {
  const ctx = {
    imageSmoothingEnabled: true, font: '', textAlign: '', textBaseline: '', fillStyle: '', _image: null,
    createImageData(w, h) { return { data: new Uint8ClampedArray(w * h * 4), width: w, height: h }; },
    putImageData(img) { captured.puts++; captured.imagedata = img; },
    drawImage(img) { ctx._image = img; },
    getImageData() { const t = ctx._image._tex; return { data: new Uint8ClampedArray(t.bytes), width: t.width, height: t.height }; },
    fillText(msg) { captured.text = String(msg); },
  };
  return ctx;
}
function makeCanvas() { const ctx = makeContext2d(); return { width: 300, height: 150, getContext() { return ctx; } }; }
const mainCanvas = makeCanvas();
const body = { clientWidth: width, clientHeight: height, onload: null, appendChild() {} };
const document = {
  body,
  createElement(tag) { return tag === 'canvas' ? makeCanvas() : { innerHTML: '' }; },
  getElementById(id) { return id === 'canvasID' ? mainCanvas : null; },
};
class Image {
  constructor() { this.onload = null; this.width = 0; this.height = 0; }
  set src(v) {
    const t = textures[path.basename(v)];
    if (!t) throw new Error('no decoded texture for ' + v);
    this._tex = { width: t.width, height: t.height, bytes: Buffer.from(t.rgba, 'base64') };
    this.width = t.width; this.height = t.height;
    queue.push([() => this.onload({ target: this }), []]);       // browsers fire onload asynchronously
  }
}
const mathShim = Object.create(Math);
mathShim.random = () => randomValue;
const sandbox = {
  document, window: { onerror: null }, Image, Math: mathShim,
  console: { log() {} },
  setTimeout(fn, _ms, ...args) { queue.push([fn, args]); return queue.length; },
  Date, JSON, Infinity, Uint8ClampedArray, Array, Object,
};
vm.createContext(sandbox);

// classic script (not ESM) so that the top-level function declarations (intersectWorld, createSphere,
// createMaterial, sampleTexture, ...) become context globals a harness can call for non-demo configs
const src = fs.readFileSync(path.join(refDir, 'main.js'), 'utf8');
const t0 = process.hrtime.bigint();
new vm.Script(src, { filename: 'main.js' }).runInContext(sandbox);
body.onload();                                                  // document.body.onload = main (main.js:17)
let ticks = 0;
while (queue.length) { const [fn, args] = queue.shift(); fn(...args); ticks++; }
const ms = Number(process.hrtime.bigint() - t0) / 1e6;

if (captured.puts !== height) throw new Error(`expected ${height} putImageData calls, saw ${captured.puts}`);
fs.writeFileSync(outPath, Buffer.from(captured.imagedata.data.buffer));
console.log(JSON.stringify({ width, height, ticks, puts: captured.puts, overlay: captured.text, ms, cores_used: 1,
                             host_cores: require('os').cpus().length, random: randomValue, node: process.version }));
