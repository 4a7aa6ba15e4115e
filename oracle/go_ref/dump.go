// dump.go — runs the UNMODIFIED reference (jamiec7919/vermeer) on a .vnf scene and writes what the parity harness needs to
// turn "parity unpinned" into a pinned oracle (tests/test_go_reference.py):
//
//	scramble.bin  the per-pixel framescramble table core.Render will draw: 6 little-endian uint64 per pixel
//	              {lensU, lensV, time, lambda, scramble0, scramble1} (core/render.go:18-23,166-176)
//	frame.float   core.FrameBuf() after -maxiter iterations: XRes*YRes*3 little-endian float32, row 0 = top
//	stats.txt     "<seconds> <rays> <shadow rays>" (core.RenderStats)
//	hits.bin      (only with -rays) one 32-byte record per input ray {t, Bu, Bv, 0 float32; ElemID, geom, NodesT, LeafsT int32},
//	              ElemID = geom = -1 on a miss; geom = index of the hit Geom's Name() in names.txt
//
// The table: core.Render fills framescramble from the package-level math/rand source. With GODEBUG=randautoseed=0 (Go >= 1.20;
// older Go always behaves so) that source is rand.NewSource(1), so a private rand.New(rand.NewSource(1)) draws the very same
// numbers first. Nothing in Init/Parse/PreRender touches math/rand (checked: only core/render.go:71,169-174 and the unused
// math/sample jitter helpers do).
//
// THIS FILE HAS NEVER BEEN COMPILED: the build image has no Go toolchain (README.md). It uses only exported reference API:
// core.Init, nodes.Parse, core.PreRender, core.Render, core.FrameBuf, core.FrameMetrics, core.TraceProbe, RenderTask.NewRay /
// NewShaderContext, Ray.Init.
//
// Build (GOPATH mode, the reference has no go.mod):  see README.md.
package main

import (
	"encoding/binary"
	"flag"
	"fmt"
	"log"
	"math"
	"math/rand"
	"os"

	_ "github.com/jamiec7919/vermeer/builtin/camera"
	_ "github.com/jamiec7919/vermeer/builtin/driver"
	_ "github.com/jamiec7919/vermeer/builtin/filter"
	_ "github.com/jamiec7919/vermeer/builtin/geom/instance"
	_ "github.com/jamiec7919/vermeer/builtin/geom/polymesh"
	_ "github.com/jamiec7919/vermeer/builtin/light"
	"github.com/jamiec7919/vermeer/builtin/scene"
	_ "github.com/jamiec7919/vermeer/builtin/shader"
	"github.com/jamiec7919/vermeer/core"
	m "github.com/jamiec7919/vermeer/math"
	"github.com/jamiec7919/vermeer/nodes"
)

var maxiter = flag.Int("maxiter", 4, "iterations to render")
var raysFile = flag.String("rays", "", "optional file of 32-byte ray records {o[3], d[3], tmax, time float32} to TraceProbe")
var anyHit = flag.Bool("anyhit", false, "trace the rays as RayTypeShadow")
var outDir = flag.String("out", ".", "output directory")

func must(err error) {
	if err != nil {
		log.Fatal(err)
	}
}

func main() {
	flag.Parse()
	if os.Getenv("GODEBUG") != "randautoseed=0" {
		log.Printf("warning: run with GODEBUG=randautoseed=0, or scramble.bin will not be the table core.Render draws")
	}
	core.Init(scene.New())
	must(nodes.Parse(flag.Arg(0)))
	must(core.PreRender())
	w, h := core.FrameMetrics()

	// the table core.Render is about to draw from the global source (seed 1), in its order (core/render.go:168-175)
	src := rand.New(rand.NewSource(1))
	f, err := os.Create(*outDir + "/scramble.bin")
	must(err)
	for i := 0; i < w*h*6; i++ {
		must(binary.Write(f, binary.LittleEndian, src.Uint64()))
	}
	f.Close()

	exit := make(chan bool)
	stats, err := core.Render(*maxiter, exit)
	must(err)
	f, err = os.Create(*outDir + "/frame.float")
	must(err)
	must(binary.Write(f, binary.LittleEndian, core.FrameBuf()))
	f.Close()
	must(os.WriteFile(*outDir+"/stats.txt", []byte(fmt.Sprintf("%v %v %v\n", stats.Duration.Seconds(), stats.RayCount, stats.ShadowRayCount)), 0644))

	if *raysFile == "" {
		return
	}
	in, err := os.ReadFile(*raysFile)
	must(err)
	n := len(in) / 32
	task := new(core.RenderTask)
	ray := task.NewRay()
	sg := task.NewShaderContext()
	names := map[string]int32{}
	var order []string
	out, err := os.Create(*outDir + "/hits.bin")
	must(err)
	ty := core.RayTypeCamera
	if *anyHit {
		ty = core.RayTypeShadow
	}
	for i := 0; i < n; i++ {
		var r [8]float32
		for k := range r {
			r[k] = math.Float32frombits(binary.LittleEndian.Uint32(in[i*32+k*4:]))
		}
		sg.Time = r[7]
		ray.Init(ty, m.Vec3{r[0], r[1], r[2]}, m.Vec3{r[3], r[4], r[5]}, r[6], 0, sg)
		rec := [4]float32{r[6], 0, 0, 0}
		ids := [4]int32{-1, -1, 0, 0}
		if core.TraceProbe(ray, sg) {
			rec = [4]float32{ray.Tclosest, sg.Bu, sg.Bv, 0}
			name := sg.Geom.(core.Node).Name()
			idx, ok := names[name]
			if !ok {
				idx = int32(len(order))
				names[name] = idx
				order = append(order, name)
			}
			ids[0], ids[1] = int32(sg.ElemID), idx
		}
		ids[2], ids[3] = int32(ray.NodesT), int32(ray.LeafsT)
		must(binary.Write(out, binary.LittleEndian, rec))
		must(binary.Write(out, binary.LittleEndian, ids))
	}
	out.Close()
	nf, err := os.Create(*outDir + "/names.txt")
	must(err)
	for _, s := range order {
		fmt.Fprintln(nf, s)
	}
	nf.Close()
}
