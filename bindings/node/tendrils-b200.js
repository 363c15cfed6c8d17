/**
 * Drop-in for the particle step of keeffEoghan/tendrils: subclasses the reference's `Tendrils`
 * (src/index.js) and routes `step()`, the flow half of `draw()`, `spawn()` and `spawnShader()` to the
 * CUDA library through the N-API addon; everything else (view rendering, buffers, colour map) is
 * inherited untouched and keeps running in WebGL.
 *
 * NOT EXECUTED IN THIS REPOSITORY (no Node in the image); the same logic is exercised through the
 * Python mirror `tendrils_b200/tendrils.py`, which the tests drive.
 *
 *   import { Tendrils } from 'tendrils/src';            // the reference
 *   import { accelerate } from './tendrils-b200';
 *   const tendrils = new (accelerate(Tendrils))(gl, options);
 */
const addon = require('./build/Release/tendrils_b200.node');

const TARGET = { state: 0, targets: 1 };
const BUF = { current: 0, previous: 1, targets: 2, flow: 3 };
// which built-in each reference shader source corresponds to (matched by the fragment source text)
const VARIANT = { direct: 0, best: 1, bright: 2, color: 3, data: 4, flow: 5 };

export const accelerate = (Tendrils, shaderKinds /* Map(shader -> {kind, variant}) */) =>
  class TendrilsB200 extends Tendrils {
    setupParticles(rootNum = this.state.rootNum, numBuffers = 2) {
      super.setupParticles(rootNum, numBuffers);          // keeps the GL objects the view renderer needs
      const [w, h] = this.particles.shape;
      this.b200 = addon.create({ particlesW: w, particlesH: h, flowW: this.viewRes[0] || 1,
        flowH: this.viewRes[1] || 1, device: 0 });
      this.hostState = new Float32Array(w*h*4);
      return this;
    }

    resize() {
      super.resize();
      if(this.b200) { addon.resizeFlow(this.b200, this.viewRes[0], this.viewRes[1]); }
      return this;
    }

    clearFlow() {
      addon.clearFlow(this.b200);
      return this;
    }

    step() {
      if(!this.timer.paused) {
        if(this.logicShader !== this.particles.logic && shaderKinds.get(this.particles.logic).kind !== 'logic') {
          throw new Error('tendrils-b200: custom logic shaders are not supported');
        }
        addon.setState(this.b200, this.state, this.viewSize);
        addon.step(this.b200, this.timer.time, this.timer.dt);
      }
      return this;
    }

    draw() {
      addon.setState(this.b200, this.state, this.viewSize);
      addon.splatFlow(this.b200, this.timer.time);
      // The untouched WebGL view renderer reads the state as textures: there is no CUDA<->WebGL interop
      // from ANGLE/headless-gl, so the state goes back through the host (SURVEY H6).  Skip when not drawing.
      addon.download(this.b200, BUF.current, this.hostState);
      this.particles.buffers[0].color[0].setPixels(this.particles.pixels /* ndarray view of hostState */);
      return this.drawView();                             // the reference's view half of draw()
    }

    /**
     * step() for apps that keep the particle state on the CPU (`particles.pixels`, src/particles.js:76-78): one pipelined
     * upload + logic pass + download (tb_step_streamed).  Asynchronous: `sync()` before the host array is read.
     */
    stepStreamed(hostState = this.particles.pixels.data, chunks = 16) {
      addon.setState(this.b200, this.state, this.viewSize);
      addon.stepStreamed(this.b200, this.timer.time, this.timer.dt, hostState, hostState, chunks);
      return this;
    }

    sync() { addon.sync(this.b200); return this; }

    spawn(spawner) {
      if(spawner === undefined) { addon.reset(this.b200); return this; }
      super.spawn(spawner);                               // fills this.particles.pixels on the CPU
      addon.upload(this.b200, BUF.current, this.particles.pixels.data);
      addon.upload(this.b200, BUF.previous, this.particles.pixels.data);
      return this;
    }

    spawnShader(shader, update, buffer) {
      this.timer.tick();                                  // src/index.js:433
      const { kind, variant } = shaderKinds.get(shader) || {};
      const target = ((buffer === this.targets)? TARGET.targets : TARGET.state);
      const u = ((typeof update === 'function')? update({ ...this.state, time: this.timer.time,
        viewSize: this.viewSize, viewRes: this.viewRes }) : { ...this.state, ...update });
      addon.setState(this.b200, this.state, this.viewSize);
      if(kind === 'init') { addon.spawnInit(this.b200, target); }
      else if(kind === 'ball') { addon.spawnBall(this.b200, u.radius, u.speed, target); }
      else if(kind === 'pixels') {
        const source = ((u.spawnDataIs === 'flow')? 1 : ((u.spawnDataIs === 'particles')? 2 : 0));
        if(source === 0) { addon.setSpawnImage(this.b200, u.spawnPixels, u.spawnShape[0], u.spawnShape[1]); }
        addon.spawnPixels(this.b200, u, VARIANT[variant], source, this.timer.time, target);
      }
      else { throw new Error('tendrils-b200: custom spawn shaders are not supported'); }
      return this;
    }
  };

/**
 * The inputs the app draws into the flow FBO after the particles (src/demo.main.js:1107-1159), routed to the
 * library's grid instead:
 *   accelerateFlowLine(FlowLine)      -- `flowLine.update().draw(tendrils)` hands the attribute arrays that
 *                                        Line.update() filled (src/geom/line/index.js:73-117) to tb_flow_line;
 *   accelerateOpticalFlow(OpticalFlow) -- `opticalFlow.update(u).render(tendrils, view, last, [w, h])` hands the
 *                                        two RGBA8 frames (what setPixels uploaded) to tb_optical_flow.
 */
export const accelerateFlowLine = (FlowLine) =>
  class FlowLineB200 extends FlowLine {
    draw(tendrils) {
      const { line } = this;
      if(line.path.length > 0) {
        const a = line.attributes;
        addon.flowLine(tendrils.b200, line.uniforms, a.position.data, a.normal.data, a.miter.data,
          a.previous.data, a.time.data, a.dt.data);
      }
      return this;
    }
  };

export const accelerateOpticalFlow = (OpticalFlow) =>
  class OpticalFlowB200 extends OpticalFlow {
    render(tendrils, view, last, [w, h]) {
      addon.opticalFlow(tendrils.b200, this.uniforms, view, last, w, h);
      return this;
    }
  };

export default accelerate;
