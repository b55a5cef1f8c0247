// dump -- headless driver of the REFERENCE package github.com/TheFellow/fluid/pkg/fluid:
// preset -> N x Simulate -> raw little-endian float32 dumps of U, V, M, p.
//
// It restates what main/ does around the solver and nothing else: NewGame + resetToPreset +
// applyWallSettings (main/main.go:200-229, 747-793, 840-849), the per-frame jet (main/main.go:236-241)
// and applySources (main/main.go:474-486), dt = float32(1.0/120.0) (main/main.go:243).  Sizes other than
// 300x251 scale the jet span and the obstacle radius like fluid_b200/presets.py (round(v*H/251)).
//
// Usage (inside a checkout of the reference, see tests/golden/verify_with_go.sh):
//   go run ./cmd/dump -preset karman -w 160 -h 80 -bfecc -confinement 0.1 -steps 1,10,100 -out /tmp/dump
// writes /tmp/dump/<field>_<step>.f32 (NumX*NumY float32, index i*NumY+j) and meta.txt.
//
// This file is the pin the oracle is waiting for: the image this repository is built in has no Go
// toolchain, so the goldens under tests/golden come from the C restatement; wherever `go` exists,
// verify_with_go.sh compares them with THIS program's output bit for bit.
package main

import (
	"encoding/binary"
	"flag"
	"fmt"
	"math"
	"os"
	"path/filepath"
	"strconv"
	"strings"

	"github.com/TheFellow/fluid/pkg/fluid"
)

func scaled(at251, height int) int {
	return int(math.Round(float64(at251) * float64(height) / 251.0))
}

func walls(f *fluid.Fluid, top, bottom, left, right bool) { // applyWallSettings, main/main.go:840-849
	for i := 0; i < f.NumX; i++ {
		f.SetSolid(i, 0, bottom)
		f.SetSolid(i, f.NumY-1, top)
	}
	for j := 0; j < f.NumY; j++ {
		f.SetSolid(0, j, left)
		f.SetSolid(f.NumX-1, j, right)
	}
}

func dump(dir, name string, step int, a []float32) {
	buf := make([]byte, 4*len(a))
	for k, v := range a {
		binary.LittleEndian.PutUint32(buf[4*k:], math.Float32bits(v))
	}
	if err := os.WriteFile(filepath.Join(dir, fmt.Sprintf("%s_%d.f32", name, step)), buf, 0o644); err != nil {
		panic(err)
	}
}

func main() {
	preset := flag.String("preset", "jet", "jet | cavity | karman")
	w := flag.Int("w", 300, "interior width")
	h := flag.Int("h", 251, "interior height")
	bfecc := flag.Bool("bfecc", false, "UseBFECC")
	conf := flag.Float64("confinement", 0, "Confinement")
	stepsArg := flag.String("steps", "1,10,100", "snapshot steps, ascending")
	out := flag.String("out", "dump", "output directory")
	flag.Parse()

	f := fluid.New(1000, *w, *h, 1.0/100.0) // main/main.go:201
	for i := 0; i < f.NumX; i++ {          // main/main.go:222-226
		for j := 0; j < f.NumY; j++ {
			f.SetSolid(i, j, false)
		}
	}
	f.UseBFECC = *bfecc
	f.Confinement = float32(*conf)

	span := 100
	radius := 8
	if *h != 251 {
		span = scaled(100, *h)
		radius = scaled(8, *h)
		if radius < 1 {
			radius = 1
		}
	}
	jet := false
	type source struct {
		i, j int
		u, v float32
	}
	var sources []source
	switch *preset {
	case "jet": // main/main.go:760-765
		walls(f, true, true, true, false)
		jet = true
	case "cavity": // main/main.go:767-779
		for i := 2; i < f.NumX-2; i++ {
			sources = append(sources, source{i, f.NumY - 2, 3.0, 0})
		}
		walls(f, true, true, true, true)
	case "karman": // main/main.go:781-790, with the jet on
		f.SetCircularObstacle(f.NumX/4, f.NumY/2, radius)
		walls(f, true, true, true, false)
		jet = true
	default:
		panic("unknown preset " + *preset)
	}

	if err := os.MkdirAll(*out, 0o755); err != nil {
		panic(err)
	}
	dt := float32(1.0 / 120.0) // main/main.go:243 at speed 1
	done := 0
	for _, tok := range strings.Split(*stepsArg, ",") {
		target, err := strconv.Atoi(strings.TrimSpace(tok))
		if err != nil {
			panic(err)
		}
		for ; done < target; done++ {
			for _, s := range sources { // applySources, main/main.go:474-480
				if !f.IsSolid(s.i, s.j) {
					f.SetVelocity(s.i, s.j, s.u, s.v)
					f.AddSmoke(s.i, s.j, 0.5)
				}
			}
			if jet { // main/main.go:236-241
				for j := *h/2 - span; j < *h/2+span; j++ {
					f.SetVelocity(1, j, 4.0, 0)
					f.AddSmoke(1, j, 1.0)
				}
			}
			f.Simulate(dt)
		}
		dump(*out, "U", target, f.U)
		dump(*out, "V", target, f.V)
		dump(*out, "M", target, f.M)
		pf := f.Pressure() // p is unexported: read it back through the view (pressure.go:5-24)
		pv := make([]float32, f.NumX*f.NumY)
		for i := 0; i < f.NumX; i++ {
			for j := 0; j < f.NumY; j++ {
				pv[i*f.NumY+j], _ = pf.Value(i, j)
			}
		}
		dump(*out, "p", target, pv)
	}
	meta := fmt.Sprintf("preset=%s NumX=%d NumY=%d bfecc=%v confinement=%g steps=%s\n", *preset, f.NumX, f.NumY, *bfecc, *conf, *stepsArg)
	_ = os.WriteFile(filepath.Join(*out, "meta.txt"), []byte(meta), 0o644)
}
