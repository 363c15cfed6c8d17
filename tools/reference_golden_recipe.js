// UNVERIFIED RECIPE (cannot run in this image: no Node, no GL) -- how to mint reference-held golden vectors for the hot path on a
// machine with Node.js and headless-gl (`npm i gl`), from a checkout of keeffEoghan/tendrils at /root/reference.  See SURVEY.md 8(c)
// and spec/PARITY.md: until this has been run, RASTER-1, B2, T1/T2 and TSIN-1 are decisions, and parity is "unpinned".
//
//   node tools/reference_golden_recipe.js R G steps out.json
//
// writes the particle state (RGBA32F, R x R) and the flow grid (G x G) after `steps` x (tick, step, draw) from a ball spawn, as JSON
// arrays of floats, for tests/ to compare with oracle/tendrils_oracle.c on the same inputs.
const [R, G, steps, out] = [parseInt(process.argv[2] || '64'), parseInt(process.argv[3] || '32'), parseInt(process.argv[4] || '8'), process.argv[5] || 'golden.json'];
global.location = { href: 'http://x/' };                                  // src/utils/index.js:20-23 reads it
const gl = require('gl')(G, G, { preserveDrawingBuffer: true });
if (!gl.getExtension('OES_texture_float')) throw new Error('OES_texture_float is required');
const { Tendrils } = require('/root/reference/docs/js/index.js');
const spawnBall = require('/root/reference/src/spawn/ball').default;
const t = new Tendrils(gl, {});                                           // fixed-step timer is the default (src/index.js:67)
t.setup(R); t.resize();
spawnBall(gl, { uniforms: { radius: 0.3, speed: 0.005 } }).spawn(t);
for (let k = 0; k < steps; ++k) { t.timer.tick(); t.step(); t.draw(); }
const read = (fbo, w, h) => { fbo.bind(); const a = new Float32Array(4 * w * h); gl.readPixels(0, 0, w, h, gl.RGBA, gl.FLOAT, a); return Array.from(a); };
require('fs').writeFileSync(out, JSON.stringify({ R, G, steps, state: read(t.particles.buffers[0], R, R), flow: read(t.flow, G, G) }));
