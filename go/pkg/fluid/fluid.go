// Package fluid is a drop-in replacement for github.com/TheFellow/fluid/pkg/fluid whose
// per-step solver runs on an NVIDIA B200 through libfluidb200.so (include/fluidb200.h).
//
// Every exported name of the reference package is kept with the same meaning:
// New, (*Fluid).Simulate, the edits of walls.go, the views of pressure.go / smoke.go /
// velocity.go / fluid.go:799-891, ScalarField, VectorField and the package variable
// Relaxation.  main/ compiles and runs unchanged against this package.
//
// NOTE: this image has no Go toolchain, so this file has not been compiled here.  It is
// deliberately thin and declarative: all logic that could be tested lives behind the C ABI
// and is exercised through the identical Python binding (fluid_b200/fluid.py).
//
// Build: CGO_CFLAGS=-I<repo>/include CGO_LDFLAGS="-L<repo>/fluid_b200 -lfluidb200" go build ./...
package fluid

/*
#cgo LDFLAGS: -lfluidb200
#include <stdlib.h>
#include "fluidb200.h"
*/
import "C"

import (
	"fmt"
	"math"
	"runtime"
	"unsafe"
)

// Relaxation mirrors the reference's package variable (fluid.go:7-9).
var Relaxation float32 = 1.9

// Solver selects the ordering of the pressure solve (not in the reference).
type Solver int32

const (
	SolverExact            Solver = C.FB_SOLVER_EXACT             // lexicographic, bit-identical to the reference
	SolverRedBlack         Solver = C.FB_SOLVER_REDBLACK          // red-black on the face velocities
	SolverRedBlackPressure Solver = C.FB_SOLVER_REDBLACK_PRESSURE // red-black in pressure form (fastest)
)

// Fluid keeps the exported fields of the reference's struct (fluid.go:11-40).
// U, V, S, M are backed by pinned host mirrors owned by the C library; they are refreshed
// by Sync() and by the view methods, and are authoritative only in Compat mode.
type Fluid struct {
	NumX, NumY int
	U, V       []float32
	S          []float32
	M          []float32

	Confinement        float32
	ViscosityDiffusion float32
	PressureDamping    float32
	TurbulenceStrength float32
	SmokeAdvection     float32
	UseMultigrid       bool
	MultigridLevels    int
	UseBFECC           bool

	// Not in the reference.
	// Solver: New() starts with SolverRedBlackPressure, the ordering the published throughput refers to
	// (red-black, 8 iterations, same residual as the reference's 8 lexicographic sweeps; fields differ
	// from the reference's by the size of the solver residual).  Set SolverExact for the reference's own
	// sweep order, bit for bit, at ~13x the step time on large grids.  NewCompat() starts with SolverExact.
	Solver Solver
	Compat bool // upload U,V,S,M before and download after every call (white-box tests)

	h        *C.fb_handle
	density  float32
	spacing  float32
	numCells int
	pending  []C.fb_edit_cmd // edits queued since the last flush
	solidOK  bool            // the S mirror reflects every edit issued so far
	uvOK     bool            // the U, V mirrors reflect the device (SampleVelocity is served from them)
	frame    int             // which pinned frame buffer the view in flight targets (BeginSmoke)
	frameBuf []float32
	// display frame of BeginSmokeFrame: C-allocated bytes (the asynchronous copy outlives the cgo call, so it must not
	// target Go memory) and their shape
	shotC                           unsafe.Pointer
	shot                            []byte
	shotLines, shotCols, shotStride int
}

func check(h *C.fb_handle, st C.int) {
	if st != C.FB_OK {
		panic(fmt.Sprintf("fluidb200: status %d: %s", int(st), C.GoString(C.fb_last_error(h))))
	}
}

func mirror(h *C.fb_handle, field C.int32_t) []float32 {
	var p *C.float
	var n C.size_t
	check(h, C.fb_host_mirror(h, field, &p, &n))
	return unsafe.Slice((*float32)(unsafe.Pointer(p)), int(n))
}

// New replaces fluid.New (fluid.go:42-68).  Compat handles are created with
// FB_FLAG_EXACT_SHADOW because white-box callers may rewrite S directly.
func New(density float32, width, height int, h float32) *Fluid {
	return newFluid(density, width, height, h, 0)
}

// NewCompat is New for callers that read and write the U, V, S, M slices directly, like the
// reference's in-package tests and benchmarks (fluid_bench_test.go:9-13).
func NewCompat(density float32, width, height int, h float32) *Fluid {
	f := newFluid(density, width, height, h, C.FB_FLAG_EXACT_SHADOW)
	f.Compat = true
	return f
}

func newFluid(density float32, width, height int, h float32, flags C.int32_t) *Fluid {
	cfg := C.fb_config{width: C.int32_t(width), height: C.int32_t(height), density: C.float(density),
		h: C.float(h), device: 0, rank: 0, nranks: 1, flags: flags}
	var handle *C.fb_handle
	if st := C.fb_create(&cfg, &handle); st != C.FB_OK {
		panic(fmt.Sprintf("fluidb200: fb_create failed with status %d (is a CUDA device visible?)", int(st)))
	}
	var p C.fb_params
	C.fb_default_params(&p)
	f := &Fluid{
		NumX: width + 2, NumY: height + 2, numCells: (width + 2) * (height + 2),
		Confinement: float32(p.confinement), ViscosityDiffusion: float32(p.viscosity_diffusion),
		PressureDamping: float32(p.pressure_damping), TurbulenceStrength: float32(p.turbulence_strength),
		SmokeAdvection: float32(p.smoke_advection), UseMultigrid: p.use_multigrid != 0,
		MultigridLevels: int(p.multigrid_levels), UseBFECC: p.use_bfecc != 0,
		Solver: SolverRedBlackPressure, h: handle, density: density, spacing: h,
	}
	if flags&C.FB_FLAG_EXACT_SHADOW != 0 {
		f.Solver = SolverExact
	}
	f.U, f.V = mirror(handle, C.FB_U), mirror(handle, C.FB_V)
	f.S, f.M = mirror(handle, C.FB_S), mirror(handle, C.FB_M)
	runtime.SetFinalizer(f, func(f *Fluid) {
		C.fb_destroy(f.h) // waits for the copy stream, so nothing writes the display buffer any more
		if f.shotC != nil {
			C.free(f.shotC)
		}
	})
	return f
}

// H returns the grid spacing (fluid.go:71).
func (f *Fluid) H() float32 { return f.spacing }

func b2i(b bool) C.int32_t {
	if b {
		return 1
	}
	return 0
}

// params is rebuilt for every call because callers mutate the struct fields directly
// (main/main.go:271-278 toggles Confinement and UseBFECC).
func (f *Fluid) params() C.fb_params {
	return C.fb_params{relaxation: C.float(Relaxation), confinement: C.float(f.Confinement),
		viscosity_diffusion: C.float(f.ViscosityDiffusion), pressure_damping: C.float(f.PressureDamping),
		turbulence_strength: C.float(f.TurbulenceStrength), smoke_advection: C.float(f.SmokeAdvection),
		use_multigrid: b2i(f.UseMultigrid), multigrid_levels: C.int32_t(f.MultigridLevels),
		use_bfecc: b2i(f.UseBFECC), solver: C.int32_t(f.Solver), iters: 8}
}

var mirrored = []C.int32_t{C.FB_U, C.FB_V, C.FB_S, C.FB_M}

func (f *Fluid) upload() {
	for _, fld := range mirrored {
		check(f.h, C.fb_upload(f.h, fld, (*C.float)(unsafe.Pointer(&mirror(f.h, fld)[0]))))
	}
}

// Sync refreshes the U, V, S, M mirrors from the device.
func (f *Fluid) Sync() {
	f.flush()
	for _, fld := range mirrored {
		check(f.h, C.fb_download(f.h, fld, (*C.float)(unsafe.Pointer(&mirror(f.h, fld)[0]))))
	}
	f.solidOK = true
	f.uvOK = true
}

// syncVelocity refreshes only the U, V mirrors (two planes instead of four).
func (f *Fluid) syncVelocity() {
	f.flush()
	if f.uvOK {
		return
	}
	check(f.h, C.fb_download(f.h, C.FB_U, (*C.float)(unsafe.Pointer(&f.U[0]))))
	check(f.h, C.fb_download(f.h, C.FB_V, (*C.float)(unsafe.Pointer(&f.V[0]))))
	f.uvOK = true
}

// flush sends the edits queued since the last call, in issue order, as ONE fb_edit.
// The reference's UI issues storms of tiny edits (76 406 SetSolid calls at start-up,
// main/main.go:222-226); batching keeps that to one cgo call per frame.
func (f *Fluid) flush() {
	if len(f.pending) == 0 {
		return
	}
	if f.Compat {
		f.upload()
	}
	check(f.h, C.fb_edit(f.h, &f.pending[0], C.size_t(len(f.pending))))
	f.pending = f.pending[:0]
	f.uvOK = false
	if f.Compat {
		f.Sync()
	}
}

// Simulate replaces (*Fluid).Simulate (fluid.go:79-109).
func (f *Fluid) Simulate(dt float32) {
	f.flush()
	if f.Compat {
		f.upload()
	}
	p := f.params()
	check(f.h, C.fb_step(f.h, &p, C.float(dt), 1, nil, 0))
	f.uvOK = false
	if f.Compat {
		f.Sync()
	}
}

func (f *Fluid) checkIndex(i, j int) {
	if i < 0 || i >= f.NumX {
		panic(fmt.Sprintf("invalid x-index: %d", i)) // walls.go:7
	}
	if j < 0 || j >= f.NumY {
		panic(fmt.Sprintf("invalid y-index: %d", j)) // walls.go:10
	}
}

func (f *Fluid) push(op C.int32_t, i, j int, a, b float32) {
	f.pending = append(f.pending, C.fb_edit_cmd{op: op, i0: C.int32_t(i), j0: C.int32_t(j),
		i1: C.int32_t(i + 1), j1: C.int32_t(j + 1), a: C.float(a), b: C.float(b)})
}

// SetSolid replaces walls.go:5-49.
func (f *Fluid) SetSolid(i, j int, value bool) {
	f.checkIndex(i, j)
	v := float32(0)
	if value {
		v = 1
	}
	f.push(C.FB_EDIT_SET_SOLID, i, j, v, 0)
	f.solidOK = false
}

// IsSolid replaces walls.go:51-60 (served from the host mirror of S).
func (f *Fluid) IsSolid(i, j int) bool {
	f.checkIndex(i, j)
	if !f.solidOK {
		f.flush()
		check(f.h, C.fb_download(f.h, C.FB_S, (*C.float)(unsafe.Pointer(&f.S[0]))))
		f.solidOK = true
	}
	return f.S[i*f.NumY+j] == 0.0
}

// SetVelocity replaces walls.go:62-72.
func (f *Fluid) SetVelocity(i, j int, u, v float32) {
	f.checkIndex(i, j)
	f.push(C.FB_EDIT_SET_VELOCITY, i, j, u, v)
}

// AddSmoke replaces walls.go:74-83.
func (f *Fluid) AddSmoke(i, j int, smoke float32) {
	f.checkIndex(i, j)
	f.push(C.FB_EDIT_ADD_SMOKE, i, j, smoke, 0)
}

// Reset replaces walls.go:85-93.
func (f *Fluid) Reset() { f.pending = append(f.pending, C.fb_edit_cmd{op: C.FB_EDIT_RESET}) }

// ApplyForce replaces fluid.go:761-771 (the ring / solid test happens on the device).
func (f *Fluid) ApplyForce(i, j int, fx, fy float32) { f.push(C.FB_EDIT_APPLY_FORCE, i, j, fx, fy) }

// ApplyForceRadius replaces fluid.go:774-796; the brush loop and Go's own math.Exp stay here,
// so the weights are bit-identical to the reference's.
func (f *Fluid) ApplyForceRadius(cx, cy int, fx, fy float32, radius int) {
	if radius <= 0 {
		f.ApplyForce(cx, cy, fx, fy)
		return
	}
	r2 := float32(radius * radius)
	for i := cx - radius; i <= cx+radius; i++ {
		for j := cy - radius; j <= cy+radius; j++ {
			if i < 1 || i >= f.NumX-1 || j < 1 || j >= f.NumY-1 {
				continue
			}
			dx, dy := float32(i-cx), float32(j-cy)
			dist2 := dx*dx + dy*dy
			if dist2 > r2 {
				continue
			}
			weight := float32(math.Exp(float64(-3.0 * dist2 / r2)))
			f.ApplyForce(i, j, fx*weight, fy*weight)
		}
	}
}

// SetCircularObstacle replaces fluid.go:894-907.
func (f *Fluid) SetCircularObstacle(cx, cy, radius int) {
	f.pending = append(f.pending, C.fb_edit_cmd{op: C.FB_EDIT_CIRCLE_OBSTACLE, i0: C.int32_t(cx),
		j0: C.int32_t(cy), i1: C.int32_t(radius)})
	f.solidOK = false
}

// sampleMirror is sampleField (fluid.go:357-398) on a host mirror: the same float32 operations in the
// same order (Go on amd64 never fuses them), dx / dy the staggering offsets of the field.
func (f *Fluid) sampleMirror(data []float32, x, y, dx, dy float32) float32 {
	n := f.NumY
	h := f.spacing
	h1 := float32(1.0 / h)
	x = max(min(x, float32(f.NumX)*h), h)
	y = max(min(y, float32(f.NumY)*h), h)
	x0 := min(int(math.Floor(float64((x-dx)*h1))), f.NumX-1)
	tx := ((x - dx) - float32(x0)*h) * h1
	x1 := min(x0+1, f.NumX-1)
	y0 := min(int(math.Floor(float64((y-dy)*h1))), f.NumY-1)
	ty := ((y - dy) - float32(y0)*h) * h1
	y1 := min(y0+1, f.NumY-1)
	sx := 1.0 - tx
	sy := 1.0 - ty
	return sx*sy*data[x0*n+y0] + tx*sy*data[x1*n+y0] + tx*ty*data[x1*n+y1] + sx*ty*data[x0*n+y1]
}

// SampleVelocity replaces fluid.go:799-803.  It is served from the U, V host mirrors, which are refreshed
// at most once per Simulate (two plane downloads on the first sample after a step), so main/'s particle loop
// (main/main.go:512-546: two samples per particle per frame) costs no cgo call and no kernel launch per point.
func (f *Fluid) SampleVelocity(x, y float32) (float32, float32) {
	f.syncVelocity()
	h2 := float32(0.5 * f.spacing)
	return f.sampleMirror(f.U, x, y, 0, h2), f.sampleMirror(f.V, x, y, h2, 0)
}

// SampleVelocities samples n points at once ON THE DEVICE (fb_sample_velocity): xy and the result are [n][2]
// flattened.  For large batches that should not wait for a two-plane download.
func (f *Fluid) SampleVelocities(xy []float32) []float32 {
	f.flush()
	uv := make([]float32, len(xy))
	if len(xy) == 0 {
		return uv
	}
	check(f.h, C.fb_sample_velocity(f.h, C.size_t(len(xy)/2), (*C.float)(unsafe.Pointer(&xy[0])),
		(*C.float)(unsafe.Pointer(&uv[0]))))
	return uv
}

func (f *Fluid) view(kind C.int32_t) ScalarField {
	f.flush()
	if f.Compat {
		f.upload()
	}
	vals := make([]float32, f.numCells)
	var mn, mx C.float
	check(f.h, C.fb_view(f.h, kind, (*C.float)(unsafe.Pointer(&vals[0])), &mn, &mx))
	return ScalarField{NumX: f.NumX, NumY: f.NumY, values: vals, MinValue: float32(mn), MaxValue: float32(mx)}
}

// Smoke replaces smoke.go:5-24.
func (f *Fluid) Smoke() ScalarField { return f.view(C.FB_VIEW_SMOKE) }

// BeginSmoke / EndSmoke are the pipelined form of Smoke() for a frame loop (fb_view_begin /
// fb_view_end): BeginSmoke queues the view behind the Simulate calls already made and returns
// at once; the frame travels to the host while the next Simulate runs; EndSmoke waits for it.
// The two frame buffers are the library's pinned mirrors of M and newM (C-owned memory, so the
// asynchronous copy never touches Go memory); a frame stays valid until the next-but-one BeginSmoke.
//
//	f.Simulate(dt); f.BeginSmoke()
//	for { f.Simulate(dt); frame := f.EndSmoke(); f.BeginSmoke(); draw(frame) }
func (f *Fluid) BeginSmoke() {
	f.flush()
	f.frame ^= 1
	field := C.int32_t(C.FB_M)
	if f.frame == 1 {
		field = C.int32_t(C.FB_NEWM)
	}
	var ptr *C.float
	var n C.size_t
	check(f.h, C.fb_host_mirror(f.h, field, &ptr, &n))
	f.frameBuf = unsafe.Slice((*float32)(unsafe.Pointer(ptr)), int(n))
	check(f.h, C.fb_view_begin(f.h, C.FB_VIEW_SMOKE, ptr))
}

func (f *Fluid) EndSmoke() ScalarField {
	var mn, mx C.float
	check(f.h, C.fb_view_end(f.h, &mn, &mx))
	return ScalarField{NumX: f.NumX, NumY: f.NumY, values: f.frameBuf, MinValue: float32(mn), MaxValue: float32(mx)}
}

// DisplayFrame is what BeginSmokeFrame / EndSmokeFrame deliver: every Stride-th cell of every Stride-th line of the
// Smoke() view as one byte, quantised on the device against the FULL field's MinValue / MaxValue
// (byte = round(255 (v - min) / (max - min))); Pix[i*Cols+j] shows cell (i*Stride, j*Stride).  For windows smaller
// than the grid (main/ draws one pixel per cell): 1/(4 Stride^2) of the bytes of EndSmoke's float field.
type DisplayFrame struct {
	Lines, Cols, Stride int
	Pix                 []byte
	MinValue, MaxValue  float32
}

// BeginSmokeFrame is BeginSmoke for a display frame (fb_view_u8_begin).  The bytes land in a C-allocated buffer: the
// copy completes after the cgo call returns (at EndSmokeFrame), and Go memory may only be lent to C for the
// duration of a call.  EndSmokeFrame's Pix aliases that buffer until the next BeginSmokeFrame.
func (f *Fluid) BeginSmokeFrame(stride int) {
	f.flush()
	if stride < 1 {
		panic("fluid: stride must be >= 1")
	}
	lines, cols := (f.NumX+stride-1)/stride, (f.NumY+stride-1)/stride
	if len(f.shot) != lines*cols {
		if f.shotC != nil {
			C.free(f.shotC)
		}
		f.shotC = C.malloc(C.size_t(lines * cols))
		f.shot = unsafe.Slice((*byte)(f.shotC), lines*cols)
	}
	var l, c C.int32_t
	check(f.h, C.fb_view_u8_begin(f.h, C.FB_VIEW_SMOKE, C.int32_t(stride), (*C.uint8_t)(f.shotC), &l, &c))
	f.shotLines, f.shotCols, f.shotStride = int(l), int(c), stride
}

func (f *Fluid) EndSmokeFrame() DisplayFrame {
	var mn, mx C.float
	check(f.h, C.fb_view_end(f.h, &mn, &mx))
	return DisplayFrame{Lines: f.shotLines, Cols: f.shotCols, Stride: f.shotStride, Pix: f.shot,
		MinValue: float32(mn), MaxValue: float32(mx)}
}

// ---- the frame loop either side of Simulate: main/'s Draw pixel pass and advectParticles ------

// VizKind selects the view Render draws (main/main.go: VizSmoke, VizPressure, VizVelMag, VizVorticity).
type VizKind int32

const (
	VizSmoke     VizKind = C.FB_VIEW_SMOKE
	VizPressure  VizKind = C.FB_VIEW_PRESSURE
	VizVelMag    VizKind = C.FB_VIEW_VELOCITY_MAGNITUDE
	VizVorticity VizKind = C.FB_VIEW_VORTICITY
)

// Render replaces drawScalarField / drawVorticityField plus the solid overlay of Draw
// (main/main.go:550-574, 620-652; main/colors.go:8-84): pix is the Pix slice of an
// image.RGBA of NumX x NumY pixels (Stride 4*NumX); it is filled exactly as those loops fill it.
// The colormap runs on the device, so a frame costs one 4-byte-per-cell transfer and no host pass.
func (f *Fluid) Render(kind VizKind, pix []uint8) (minValue, maxValue float32) {
	f.flush()
	if len(pix) < 4*f.numCells {
		panic("fluid: Render needs 4*NumX*NumY bytes")
	}
	var mn, mx C.float
	check(f.h, C.fb_render(f.h, C.int32_t(kind), (*C.uint8_t)(unsafe.Pointer(&pix[0])), nil, &mn, &mx))
	return float32(mn), float32(mx)
}

// Particle is main/main.go:139-144 with the memory layout of C.fb_particle.
type Particle struct {
	X, Y    float32
	R, G, B uint8
	_       uint8
	Age     float32
	MaxAge  float32
}

// AdvectParticles replaces the body of advectParticles (main/main.go:512-546): the slice is
// updated in place and the survivors, in their original order, are returned.
func (f *Fluid) AdvectParticles(ps []Particle, dt float32) []Particle {
	f.flush()
	if len(ps) == 0 {
		return ps
	}
	var alive C.size_t
	check(f.h, C.fb_advect_particles(f.h, (*C.fb_particle)(unsafe.Pointer(&ps[0])), C.size_t(len(ps)), C.float(dt), &alive))
	return ps[:int(alive)]
}

// Pressure replaces pressure.go:5-24.
func (f *Fluid) Pressure() ScalarField { return f.view(C.FB_VIEW_PRESSURE) }

// VelocityMagnitude replaces fluid.go:841-873.
func (f *Fluid) VelocityMagnitude() ScalarField { return f.view(C.FB_VIEW_VELOCITY_MAGNITUDE) }

// Vorticity replaces fluid.go:806-838.
func (f *Fluid) Vorticity() ScalarField { return f.view(C.FB_VIEW_VORTICITY) }

// Velocity replaces velocity.go:3-18 (the reference aliases f.U, f.V; here the two mirrors are refreshed first).
func (f *Fluid) Velocity() VectorField {
	f.syncVelocity()
	return VectorField{NumX: f.NumX, NumY: f.NumY, valuesU: f.U, valuesV: f.V}
}

// MaxDivergence replaces fluid.go:876-891.
func (f *Fluid) MaxDivergence() float32 {
	f.flush()
	var out C.float
	check(f.h, C.fb_reduce(f.h, C.FB_REDUCE_MAX_DIVERGENCE, &out))
	return float32(out)
}

// GetAdaptiveTimeStep replaces fluid.go:529-557 (the max reduction runs on the device).
func (f *Fluid) GetAdaptiveTimeStep(basedt float32) float32 {
	f.flush()
	var mv C.float
	check(f.h, C.fb_reduce(f.h, C.FB_REDUCE_MAX_ABS_VELOCITY, &mv))
	maxVel := float32(mv)
	if maxVel == 0 {
		return basedt
	}
	adaptivedt := float32(0.8) * f.spacing / maxVel
	return max(min(adaptivedt, basedt*2.0), basedt*0.1)
}
