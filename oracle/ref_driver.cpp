// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product.
//
// Driver around the UNMODIFIED reference (stark + symx + tmcd, built from /root/reference by
// oracle/Makefile.ref into oracle/_ref/).  Two jobs:
//   1. `--dump DIR`: run a deterministic scene for a few steps through the reference's public API and
//      write every input and every stage output of ONE Newton iteration (SURVEY.md section 8(c) parity
//      protocol): per-potential connectivity + bound arrays, per-element [E | grad | Hessian] outputs,
//      global E / grad, the assembled 3x3-BCSR (rows/cols/float vals), the PCG solution, the six
//      proximity lists, the edge-triangle intersection list and the PD-projected element Hessians.
//      tests/golden/make_golden.py packs these into the committed .npz fixtures.
//   2. `--bench`: time `newton->solve()` over K time steps and print Newton-iterations/s as one JSON line
//      (the CPU baseline / `bench.py --impl reference` arm).
//
// The `#define private public` below only widens access for *reading* internal state (contact tables,
// compiled potentials); no reference source is modified or copied.
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <omp.h>
#include <Eigen/Dense>
#include <Eigen/Sparse>

#define private public
#define protected public
#include <stark>
#undef private
#undef protected

namespace fs = std::filesystem;

// ---------------------------------------------------------------------------------------------------
// tiny dump helpers
// ---------------------------------------------------------------------------------------------------
struct Dumper
{
	std::string dir;
	std::ostringstream meta;
	bool first = true;

	explicit Dumper(const std::string& d) : dir(d) { fs::create_directories(d); meta << "{\n"; }
	void key(const std::string& k) { if (!first) meta << ",\n"; first = false; meta << "  \"" << k << "\": "; }
	void num(const std::string& k, double v) { key(k); meta << std::setprecision(17) << v; }
	void integer(const std::string& k, long long v) { key(k); meta << v; }
	void str(const std::string& k, const std::string& v) { key(k); meta << "\"" << v << "\""; }
	void raw(const std::string& k, const std::string& json) { key(k); meta << json; }
	template<typename T> void bin(const std::string& name, const T* p, size_t n)
	{
		std::ofstream f(dir + "/" + name + ".bin", std::ios::binary);
		if (n > 0) f.write(reinterpret_cast<const char*>(p), sizeof(T) * n);
	}
	void finish() { meta << "\n}\n"; std::ofstream f(dir + "/meta.json"); f << meta.str(); }
};

struct Args
{
	std::string scene = "tetdrop";
	int n = 4;           // grid subdivisions (per axis, tetdrop) / nx (tetbar)
	int ny = -1, nz = -1;
	int steps = 3;       // time steps to run before the dump / to time in bench mode
	int warmup = 1;      // bench: untimed time steps before timing (first step includes JIT/init)
	int threads = -1;
	std::string dump = "";
	bool bench = false;
	double drop = 0.003; // initial gap between body and floor contact surfaces' zero plane
	double dt = 0.01;
	std::string codegen = "";
	bool verbose = false;
	double vz = 0.0;     // initial downward speed (to reach contact quickly in small fixtures)
	bool llt = false;
	double inject = 1.0; // dump: injected DoF state v1 := inject * v0 + small deterministic noise
	std::string trace = ""; // trace mode: one JSON line per time step (solver statistics + state) into this file
	std::string ops = "";   // dump: name of a (user) potential whose symx operation sequences are dumped as well
};

static Args parse(int argc, char** argv)
{
	Args a;
	for (int i = 1; i < argc; i++) {
		std::string s = argv[i];
		auto next = [&]() { if (i + 1 >= argc) { std::cerr << "missing value for " << s << "\n"; exit(2); } return std::string(argv[++i]); };
		if (s == "--scene") a.scene = next();
		else if (s == "--n") a.n = std::stoi(next());
		else if (s == "--ops") a.ops = next();
		else if (s == "--ny") a.ny = std::stoi(next());
		else if (s == "--nz") a.nz = std::stoi(next());
		else if (s == "--steps") a.steps = std::stoi(next());
		else if (s == "--warmup") a.warmup = std::stoi(next());
		else if (s == "--threads") a.threads = std::stoi(next());
		else if (s == "--dump") a.dump = next();
		else if (s == "--bench") a.bench = true;
		else if (s == "--drop") a.drop = std::stod(next());
		else if (s == "--dt") a.dt = std::stod(next());
		else if (s == "--vz") a.vz = std::stod(next());
		else if (s == "--codegen") a.codegen = next();
		else if (s == "--verbose") a.verbose = true;
		else if (s == "--llt") a.llt = true;
		else if (s == "--inject") a.inject = std::stod(next());
		else if (s == "--trace") a.trace = next();
		else { std::cerr << "unknown arg " << s << "\n"; exit(2); }
	}
	return a;
}

// ---------------------------------------------------------------------------------------------------
// scenes (BASELINE.json configs, built only through the reference's public API)
// ---------------------------------------------------------------------------------------------------
struct Scene
{
	std::unique_ptr<stark::Simulation> sim;
	std::function<void()> per_step = nullptr;
	std::shared_ptr<void> keep = nullptr;   // data a user potential's lambda binds by reference
};

static stark::Settings base_settings(const Args& a, const std::string& name)
{
	stark::Settings settings;
	settings.output.simulation_name = name;
	settings.output.output_directory = "/tmp/stark_ref_out/" + name;
	if (!a.codegen.empty()) settings.output.codegen_directory = a.codegen;
	settings.output.enable_frame_writes = false;
	settings.output.enable_output = a.verbose;
	settings.output.console_verbosity = a.verbose ? symx::Verbosity::Full : symx::Verbosity::Minimal;
	settings.output.file_verbosity = symx::Verbosity::Minimal;
	settings.simulation.max_time_step_size = a.dt;
	if (a.threads > 0) settings.execution.n_threads = a.threads;
	else settings.execution.n_threads = omp_get_max_threads();
	if (a.llt) settings.newton.linear_solver = symx::LinearSolver::DirectLLT;
	return settings;
}

// C2: n^3 tet grid (Soft_Rubber) dropped onto a fixed rigid floor box, IPC contact + friction.
static Scene scene_tetdrop(const Args& a)
{
	Scene sc;
	stark::Settings settings = base_settings(a, "tetdrop");
	sc.sim = std::make_unique<stark::Simulation>(settings);
	auto& sim = *sc.sim;

	stark::EnergyFrictionalContact::GlobalParams cp;
	cp.default_contact_thickness = 0.001;
	cp.min_contact_stiffness = 1e8;
	sim.interactions->contact->set_global_params(cp);

	auto material = stark::Volume::Params::Soft_Rubber();
	auto [V, T, H] = sim.presets->deformables->add_volume_grid("body", { 1.0, 1.0, 1.0 }, { a.n, a.n, a.n }, material);
	H.point_set.add_displacement({ 0.0, 0.0, 0.5 + a.drop });
	if (a.vz != 0.0) {
		for (int i = 0; i < H.point_set.size(); i++) H.point_set.set_velocity(i, { 0.0, 0.0, -a.vz });
	}

	auto [Vf, Cf, floor] = sim.presets->rigidbodies->add_box("floor", 1.0, { 4.0, 4.0, 0.1 });
	floor.rigidbody.set_translation({ 0.0, 0.0, -0.05 });
	sim.rigidbodies->add_constraint_fix(floor.rigidbody);
	sim.interactions->contact->set_friction(floor.contact, H.contact, 0.5);
	return sc;
}

// A USER potential (examples/main.cpp:666-690, EnergyMagneticAttraction): the tet drop scene plus an attraction -k / |x1 - m| of
// every vertex to a fixed magnet, added through the public GlobalPotential::add_potential with a lambda -- no kernel of that
// name exists on the GPU path, the drop-in has to generate one from the symbolic expression.
struct MagnetData { Eigen::Vector3d center = { 0.3, 0.2, 2.5 }; double force = 40.0; symx::LabelledConnectivity<1> conn{ { "point" } }; };
static Scene scene_magnet(const Args& a)
{
	Scene sc = scene_tetdrop(a);
	auto& sim = *sc.sim;
	auto data = std::make_shared<MagnetData>();
	sc.keep = data;
	stark::core::Stark& st = sim.get_stark();
	stark::PointDynamics* dyn = sim.deformables->point_sets.get();
	for (int i = 0; i < dyn->size(); i++) data->conn.push_back({ i });
	MagnetData* d = data.get();
	stark::core::Stark* pst = &st;
	st.global_potential->add_potential("EnergyMagneticAttraction", d->conn,
		[dyn, pst, d](symx::MappedWorkspace<double>& mws, symx::Element& elem)
		{
			symx::Vector v1 = mws.make_vector(dyn->v1.data, elem["point"]);
			symx::Vector x0 = mws.make_vector(dyn->x0.data, elem["point"]);
			symx::Scalar dt = mws.make_scalar(pst->dt);
			symx::Scalar k = mws.make_scalar(d->force);
			symx::Vector m = mws.make_vector(d->center);
			symx::Vector x1 = stark::time_integration(x0, v1, dt);
			symx::Vector r = x1 - m;
			return -k / r.norm();
		}
	);
	return sc;
}

// C5: tet bar, both end caps prescribed, one cap rotating 90 deg/s, no contact.
static Scene scene_tetbar(const Args& a)
{
	Scene sc;
	stark::Settings settings = base_settings(a, "tetbar");
	settings.simulation.init_frictional_contact = false;
	sc.sim = std::make_unique<stark::Simulation>(settings);
	auto& sim = *sc.sim;
	const int nx = a.n, ny = (a.ny > 0) ? a.ny : a.n, nz = (a.nz > 0) ? a.nz : 8 * a.n;
	const double h = 1.0 / 22.0;  // element size of the 22x22x172 headline bar
	const Eigen::Vector3d dim = { nx * h, ny * h, nz * h };
	auto material = stark::Volume::Params::Soft_Rubber();
	auto [V, T, H] = sim.presets->deformables->add_volume_grid("bar", dim, { nx, ny, nz }, material);
	const double hz = 0.5 * dim[2];
	auto bc0 = sim.deformables->prescribed_positions->add_inside_aabb(H.point_set, { 0.0, 0.0, -hz }, { dim[0], dim[1], 0.001 }, stark::EnergyPrescribedPositions::Params());
	auto bc1 = sim.deformables->prescribed_positions->add_inside_aabb(H.point_set, { 0.0, 0.0, hz }, { dim[0], dim[1], 0.001 }, stark::EnergyPrescribedPositions::Params());
	auto bc1p = std::make_shared<decltype(bc1)>(bc1);
	sim.add_time_event(0.0, 1e9, [bc1p, hz](double t) { bc1p->set_transformation({ 0.0, 0.0, 0.0 }, 90.0 * t, { 0.0, 0.0, 1.0 }); });
	return sc;
}

// C1 / C3: n x n cloth grid dropped over a fixed rigid box.
static Scene scene_cloth(const Args& a, bool discrete_shells, double mu)
{
	Scene sc;
	stark::Settings settings = base_settings(a, "cloth");
	sc.sim = std::make_unique<stark::Simulation>(settings);
	auto& sim = *sc.sim;
	stark::EnergyFrictionalContact::GlobalParams cp;
	// the contact thickness has to stay below the mesh spacing (0.4 m / n): 2 mm up to n = 64, 0.3 edge lengths on finer grids
	// (at n = 256 the edges are 1.56 mm long: with 2 mm every neighbouring primitive pair would be in contact at rest)
	cp.default_contact_thickness = (a.n > 64) ? 0.3 * 0.4 / a.n : 0.002;
	sim.interactions->contact->set_global_params(cp);
	auto material = stark::Surface::Params::Cotton_Fabric();
	if (discrete_shells) material.bending.flat_rest_angle = false;
	auto [V, T, cloth] = sim.presets->deformables->add_surface_grid("cloth", Eigen::Vector2d(0.4, 0.4), { a.n, a.n }, material);
	if (a.drop != 0.003) cloth.point_set.add_displacement({ 0.0, 0.0, a.drop });
	auto box = sim.presets->rigidbodies->add_box("box", 1.0, 0.08);
	box.handler.rigidbody.add_translation({ 0.0, 0.0, -0.08 });
	auto fix = sim.rigidbodies->add_constraint_fix(box.handler.rigidbody);
	if (mu > 0.0) sim.interactions->contact->set_friction(box.handler.contact, cloth.contact, mu);
	auto fixp = std::make_shared<decltype(fix)>(fix);
	sim.add_time_event(0.0, 1e9, [fixp](double t) { fixp->set_transformation({ 0.0, 0.0, -0.08 - 0.1 * std::sin(t) }, 90.0 * t, { 0.0, 0.0, 1.0 }); });
	return sc;
}

// Rigid-rigid contact + friction and the two velocity controllers: a tilted box resting on a fixed floor box, a second
// box driven along x by a linear-velocity controller and a third one spun by an angular-velocity controller.
static Scene scene_boxes(const Args& a)
{
	Scene sc;
	stark::Settings settings = base_settings(a, "boxes");
	sc.sim = std::make_unique<stark::Simulation>(settings);
	auto& sim = *sc.sim;
	stark::EnergyFrictionalContact::GlobalParams cp;
	cp.default_contact_thickness = 0.002;
	cp.min_contact_stiffness = 1e6;
	sim.interactions->contact->set_global_params(cp);

	auto [Vf, Cf, floor] = sim.presets->rigidbodies->add_box("floor", 1.0, { 2.0, 2.0, 0.1 });
	floor.rigidbody.set_translation({ 0.0, 0.0, -0.05 });
	sim.rigidbodies->add_constraint_fix(floor.rigidbody);

	auto [V1, C1, b1] = sim.presets->rigidbodies->add_box("b1", 1.0, 0.2);
	b1.rigidbody.set_rotation(7.0, { 1.0, 0.3, 0.1 });
	b1.rigidbody.set_translation({ -0.4, 0.05, 0.1 + 0.022 + a.drop });
	b1.rigidbody.set_velocity({ 0.3, 0.0, -a.vz });
	sim.interactions->contact->set_friction(floor.contact, b1.contact, 0.5);

	auto [V2, C2, b2] = sim.presets->rigidbodies->add_box("b2", 2.0, { 0.2, 0.15, 0.1 });
	b2.rigidbody.set_translation({ 0.3, -0.3, 0.05 + 0.0035 });
	sim.interactions->contact->set_friction(floor.contact, b2.contact, 0.3);
	sim.rigidbodies->add_constraint_linear_velocity(floor.rigidbody, b2.rigidbody, { 1.0, 0.2, 0.0 }, 0.25, 5.0, 0.02);

	auto [V3, C3, b3] = sim.presets->rigidbodies->add_box("b3", 0.5, 0.12);
	b3.rigidbody.set_translation({ 0.4, 0.5, 0.4 });
	sim.rigidbodies->add_constraint_angular_velocity(floor.rigidbody, b3.rigidbody, { 0.0, 0.3, 1.0 }, 2.0, 0.4, 0.05);
	// b3 leans on b2's top edge region later in the fall: rigid-rigid edge-edge pairs
	sim.interactions->contact->set_friction(b2.contact, b3.contact, 0.2);

	// two small cubes standing on a vertex (body diagonal vertical), held in place: the vertex of b4 hovers 1.4 mm off the
	// floor's top EDGE, the one of b5 1.8 mm off the floor's CORNER -> point-edge / point-point and the edge-edge derived
	// point-edge / point-point pairs (contact and friction)
	const double h = 0.05, diag = h * std::sqrt(3.0), tilt = std::acos(1.0 / std::sqrt(3.0)) * 180.0 / M_PI;
	auto [V4, C4, b4] = sim.presets->rigidbodies->add_box("b4", 0.3, 2.0 * h);
	b4.rigidbody.set_rotation(tilt, { 1.0, -1.0, 0.0 });
	b4.rigidbody.set_translation({ 1.0012, 0.3, 0.0008 + diag });
	sim.rigidbodies->add_constraint_fix(b4.rigidbody);
	sim.interactions->contact->set_friction(floor.contact, b4.contact, 0.4);
	auto [V5, C5, b5] = sim.presets->rigidbodies->add_box("b5", 0.3, 2.0 * h);
	b5.rigidbody.set_rotation(tilt, { 1.0, -1.0, 0.0 });
	b5.rigidbody.set_translation({ 1.001, 1.0012, 0.0009 + diag });
	sim.rigidbodies->add_constraint_fix(b5.rigidbody);
	sim.interactions->contact->set_friction(floor.contact, b5.contact, 0.4);
	return sc;
}

// All five attachment potentials (examples/main.cpp:268-312 scaled down): two cloth patches attached point-triangle by
// distance, explicit point-point / point-edge / edge-edge pairs, and a rigid box attached to the second patch.
static Scene scene_attach(const Args& a)
{
	Scene sc;
	stark::Settings settings = base_settings(a, "attach");
	settings.simulation.init_frictional_contact = false;
	sc.sim = std::make_unique<stark::Simulation>(settings);
	auto& sim = *sc.sim;
	const int n = a.n;
	const double d = 1.0, hd = 0.5, gap = 0.001;
	auto params = stark::Surface::Params::Cotton_Fabric();
	auto [V1, T1, H1] = sim.presets->deformables->add_surface_grid("A", { d, d }, { n, n }, params);
	auto [V2, T2, H2] = sim.presets->deformables->add_surface_grid("B", { d, d }, { n, n }, params);
	H2.point_set.add_rotation(45.0, Eigen::Vector3d::UnitZ());
	H2.point_set.add_displacement({ d, 0.0, gap });
	const double bs = 0.25;
	const stark::Mesh<3> box_mesh = stark::make_box({ bs, bs, bs });
	auto [V, C, box] = sim.presets->rigidbodies->add_box("box", 0.1, bs);
	box.rigidbody.add_translation({ 1.7, 0.0, 0.5 * bs + 2.0 * gap });

	auto ap = stark::EnergyAttachments::Params().set_tolerance(0.01);
	sim.interactions->attachments->add_by_distance(H2.point_set, H1.point_set, H2.point_set.all(), T1, 2.0 * gap, ap);
	sim.interactions->attachments->add_by_distance(box.rigidbody, H2.point_set, box_mesh.vertices, box_mesh.conn, H2.point_set.all(), 4.0 * gap, ap);
	// explicit pairs between the two patches
	const int np = (n + 1) * (n + 1);
	sim.interactions->attachments->add(H1.point_set, H2.point_set, std::vector<int>{ 0, np / 2 }, std::vector<int>{ 1, np / 3 }, ap);
	sim.interactions->attachments->add(H1.point_set, H2.point_set, std::vector<int>{ 2, 5 }, std::vector<std::array<int, 2>>{ { 0, 1 }, { 3, 4 } },
		std::vector<std::array<double, 2>>{ { 0.25, 0.75 }, { 0.6, 0.4 } }, ap);
	sim.interactions->attachments->add(H1.point_set, H2.point_set, std::vector<std::array<int, 2>>{ { 0, 1 }, { 4, 5 } }, std::vector<std::array<int, 2>>{ { 2, 3 }, { 6, 7 } },
		std::vector<std::array<double, 2>>{ { 0.5, 0.5 }, { 0.1, 0.9 } }, std::vector<std::array<double, 2>>{ { 0.3, 0.7 }, { 0.8, 0.2 } }, ap);

	sim.deformables->prescribed_positions->add_inside_aabb(H1.point_set, { -hd, -hd, 0.0 }, { 0.001, 0.001, 0.001 }, stark::EnergyPrescribedPositions::Params());
	sim.deformables->prescribed_positions->add_inside_aabb(H1.point_set, { -hd, hd, 0.0 }, { 0.001, 0.001, 0.001 }, stark::EnergyPrescribedPositions::Params());
	return sc;
}

// C4: n^3 tet grid (Soft_Rubber, bottom face prescribed) under a chain of nb rigid boxes joined by hinges (pattern of
// examples/rb_constraint_test_scenes.cpp:173-183), first box fixed; the chain swings down onto the soft body (coupled
// solve through rigid-deformable contact + friction).  nb = --ny (default 10).
static Scene scene_tetchain(const Args& a)
{
	Scene sc;
	stark::Settings settings = base_settings(a, "tetchain");
	sc.sim = std::make_unique<stark::Simulation>(settings);
	auto& sim = *sc.sim;
	stark::EnergyFrictionalContact::GlobalParams cp;
	cp.default_contact_thickness = 0.001;
	cp.min_contact_stiffness = 1e7;
	sim.interactions->contact->set_global_params(cp);
	auto material = stark::Volume::Params::Soft_Rubber();
	material.inertia.density = 50.0;   // light foam: the block keeps its shape under its own weight
	auto [V, T, H] = sim.presets->deformables->add_volume_grid("body", { 1.0, 1.0, 1.0 }, { a.n, a.n, a.n }, material);
	H.point_set.add_displacement({ 0.0, 0.0, 0.5 });
	sim.deformables->prescribed_positions->add_inside_aabb(H.point_set, { 0.0, 0.0, 0.0 }, { 1.0, 1.0, 0.001 }, stark::EnergyPrescribedPositions::Params().set_stiffness(1e7));
	const int nb = (a.ny > 0) ? a.ny : 10;
	std::vector<stark::RigidBodyHandler> bodies;
	std::vector<stark::ContactHandler> contacts;
	for (int i = 0; i < nb; i++) {
		auto [Vb, Cb, b] = sim.presets->rigidbodies->add_box("b" + std::to_string(i), 1.0, 0.08);
		b.rigidbody.set_translation({ -0.2 + 0.1 * i, 0.0, 1.0 + 0.04 + a.drop + 0.04 });
		sim.interactions->contact->set_friction(b.contact, H.contact, 0.3);
		if (i == 0) sim.rigidbodies->add_constraint_fix(b.rigidbody);
		else {
			sim.rigidbodies->add_constraint_hinge(bodies.back(), b.rigidbody, { -0.25 + 0.1 * i, 0.0, 1.0 + 0.04 + a.drop + 0.04 }, Eigen::Vector3d::UnitY());
			sim.interactions->contact->disable_collision(contacts.back(), b.contact);
		}
		bodies.push_back(b.rigidbody);
		contacts.push_back(b.contact);
	}
	return sc;
}

// Every deformable-deformable contact / friction table, the point-side rigid-deformable ones, the three
// "_Elasticity_Only" strain potentials and both rod potentials in one scene: three cloth patches in free fall whose
// boundaries and corners face each other at less than d^, a folded strip (self contact), an elasticity-only tet block and
// two rods lying on the first patch, and three small cubes standing on a vertex just outside the edge / corner of a patch.
// Everything falls together (no supports), so the relative configuration survives the steps before the dump.
static Scene scene_zoo(const Args& a)
{
	Scene sc;
	stark::Settings settings = base_settings(a, "zoo");
	sc.sim = std::make_unique<stark::Simulation>(settings);
	auto& sim = *sc.sim;
	stark::EnergyFrictionalContact::GlobalParams cp;
	cp.default_contact_thickness = 0.002;
	cp.min_contact_stiffness = 1e4;
	sim.interactions->contact->set_global_params(cp);
	const int n = a.n;
	auto full = stark::Surface::Params::Cotton_Fabric();
	auto eonly = stark::Surface::Params::Cotton_Fabric();
	eonly.strain.elasticity_only = true;
	std::vector<stark::ContactHandler> soft, all;

	// A: elasticity-only patch at z = 0
	auto [VA, TA, A] = sim.presets->deformables->add_surface_grid("A", { 1.0, 1.0 }, { n, n }, eonly);
	soft.push_back(A.contact);
	// B: slightly rotated, overlapping A's +x boundary strip from above (pt / ee interior types + boundary pe / pp types)
	auto [VB, TB, B] = sim.presets->deformables->add_surface_grid("B", { 1.0, 1.0 }, { n, n }, full);
	B.point_set.add_rotation(3.0, Eigen::Vector3d::UnitZ());
	B.point_set.add_displacement({ 0.96, 0.013, 0.0026 });
	soft.push_back(B.contact);
	// C: corner to corner with A (point-point), 1.5 mm lower
	auto [VC, TC, Cc] = sim.presets->deformables->add_surface_grid("C", { 1.0, 1.0 }, { n, n }, full);
	Cc.point_set.add_displacement({ -1.0006, -1.0009, -0.0015 });
	soft.push_back(Cc.contact);
	// D: edge to edge with A along A's -y boundary, vertices staggered (point-edge, edge-edge point-edge)
	auto [VD, TD, Dd] = sim.presets->deformables->add_surface_grid("D", { 1.0, 1.0 }, { n, n }, full);
	Dd.point_set.add_displacement({ 0.37 / n, 1.0011, 0.0012 });
	soft.push_back(Dd.contact);
	// F: a strip folded back on itself 2.4 mm above its own lower half (self contact within one mesh)
	{
		const int m = std::max(4, n);
		std::vector<Eigen::Vector3d> v;
		std::vector<std::array<int, 3>> t;
		const double w = 0.3, L = 0.6, gap = 0.0024;
		for (int j = 0; j <= 2; j++) for (int i = 0; i <= 2 * m; i++) {
			const double s = L * i / (2.0 * m);
			Eigen::Vector3d p;
			if (i <= m) p = { s, w * j / 2.0, 0.0 };
			else p = { L - s + 0.013, w * j / 2.0 + 0.011, gap };
			v.push_back(p + Eigen::Vector3d(-0.3, -2.2, 0.0));
		}
		auto id = [&](int i, int j) { return j * (2 * m + 1) + i; };
		for (int j = 0; j < 2; j++) for (int i = 0; i < 2 * m; i++) {
			t.push_back({ id(i, j), id(i + 1, j), id(i + 1, j + 1) });
			t.push_back({ id(i, j), id(i + 1, j + 1), id(i, j + 1) });
		}
		auto shells = stark::Surface::Params::Cotton_Fabric();
		shells.bending.flat_rest_angle = false;
		auto F = sim.presets->deformables->add_surface("F", v, t, shells);
		sim.interactions->contact->set_friction(F.contact, F.contact, 0.35);
		all.push_back(F.contact);
	}
	// T: elasticity-only tet block resting 2.5 mm above A
	auto vol = stark::Volume::Params::Soft_Rubber();
	vol.strain.elasticity_only = true;
	auto [VT, TT, Tt] = sim.presets->deformables->add_volume_grid("T", { 0.3, 0.3, 0.3 }, { 2, 2, 2 }, vol);
	Tt.point_set.add_rotation(11.0, Eigen::Vector3d::UnitZ());
	Tt.point_set.add_displacement({ -0.21, 0.17, 0.15 + 0.0025 });
	soft.push_back(Tt.contact);
	// two rods lying 2.2 mm above A (complete and elasticity-only strain models)
	auto rod = stark::Line::Params::Elastic_Rubberband();
	auto [VR, SR, R1] = sim.presets->deformables->add_line_as_segments("rod1", { -0.43, -0.31, 0.0022 }, { 0.21, -0.12, 0.0022 }, 7, rod);
	rod.strain.elasticity_only = true;
	auto [VR2, SR2, R2] = sim.presets->deformables->add_line_as_segments("rod2", { -0.4, -0.1, 0.0045 }, { 0.3, -0.35, 0.0045 }, 5, rod);
	soft.push_back(R1.contact);
	soft.push_back(R2.contact);
	// stretch the rods a little (strain and strain-limit terms active)
	for (int i = 0; i < R1.point_set.size(); i++) R1.point_set.set_velocity(i, { 0.9 * i, 0.0, 0.0 });
	for (int i = 0; i < R2.point_set.size(); i++) R2.point_set.set_velocity(i, { 0.2 * i, -0.1 * i, 0.0 });

	// cubes on a vertex: vertex 1.5 mm off B's +x boundary edge, 1.7 mm off B's far corner, 1.3 mm above the inside of A
	const double h = 0.05, diag = h * std::sqrt(3.0), tilt = std::acos(1.0 / std::sqrt(3.0)) * 180.0 / M_PI;
	const Eigen::AngleAxisd rotB(3.0 * M_PI / 180.0, Eigen::Vector3d::UnitZ());
	auto onB = [&](double x, double y, double z) { Eigen::Vector3d p = rotB * Eigen::Vector3d(x, y, 0.0) + Eigen::Vector3d(0.96, 0.013, 0.0026 + z); return p; };
	std::vector<Eigen::Vector3d> tips = { onB(0.5012, 0.11, 0.0009), onB(0.5011, 0.5010, 0.0008), Eigen::Vector3d(-0.33, -0.37, 0.0013) };
	for (size_t k = 0; k < tips.size(); k++) {
		auto [Vk, Ck, bk] = sim.presets->rigidbodies->add_box("cube" + std::to_string(k), 0.3, 2.0 * h);
		bk.rigidbody.set_rotation(tilt, { 1.0, -1.0, 0.0 });
		bk.rigidbody.set_translation(tips[k] + Eigen::Vector3d(0.0, 0.0, diag));
		all.push_back(bk.contact);
	}
	for (auto& c : soft) all.push_back(c);
	for (size_t i = 0; i < all.size(); i++) for (size_t j = i + 1; j < all.size(); j++) sim.interactions->contact->set_friction(all[i], all[j], 0.2 + 0.03 * ((i + j) % 5));
	// relative sliding so that the friction potentials see a tangential velocity
	for (int i = 0; i < B.point_set.size(); i++) B.point_set.set_velocity(i, { 0.05, -0.02, 0.0 });
	for (int i = 0; i < Dd.point_set.size(); i++) Dd.point_set.set_velocity(i, { -0.03, 0.0, 0.0 });
	return sc;
}

// The five rigid-body constraint potentials no other scene holds (tests/rb_constraints.cpp:106-232 patterns): slider
// (point_on_axis), distance, distance limits, angle limit and damped spring between a fixed anchor and free boxes pushed by
// external forces / torques so that every constraint is violated (and active) at the dump.
static Scene scene_joints(const Args& a)
{
	Scene sc;
	stark::Settings settings = base_settings(a, "joints");
	settings.simulation.init_frictional_contact = false;
	sc.sim = std::make_unique<stark::Simulation>(settings);
	auto& sim = *sc.sim;
	const double mass = 1.3;
	auto I = stark::inertia_tensor_box(mass, { 0.1, 0.12, 0.08 });
	auto anchor = sim.rigidbodies->add(mass, I);
	sim.rigidbodies->add_constraint_fix(anchor);
	auto mk = [&](const Eigen::Vector3d& t, double deg) {
		auto b = sim.rigidbodies->add(mass, I);
		b.set_translation(t);
		b.set_rotation(deg, { 0.3, 1.0, -0.2 });
		return b;
	};
	auto b1 = mk({ 0.2, 0.0, 0.0 }, 5.0);
	sim.rigidbodies->add_constraint_slider(anchor, b1, { 0.05, 0.01, 0.0 }, Eigen::Vector3d(0.1, 0.2, 1.0).normalized());
	b1.add_force_at_centroid({ 7.0, -3.0, 1.0 });
	auto b2 = mk({ 0.0, 0.3, 0.0 }, -8.0);
	sim.rigidbodies->add_constraint_distance(anchor, b2, { 0.01, 0.02, 0.03 }, { 0.02, 0.28, 0.01 });
	b2.add_force_at_centroid({ 1.0, 9.0, 2.0 });
	auto b3 = mk({ 0.0, -0.3, 0.1 }, 13.0);
	sim.rigidbodies->add_constraint_distance_limits(anchor, b3, { 0.0, -0.02, 0.01 }, { 0.01, -0.27, 0.08 }, 0.2, 0.2605);
	b3.add_force_at_centroid({ 0.5, -12.0, 3.0 });
	auto b4 = mk({ -0.3, 0.0, 0.0 }, 0.0);
	sim.rigidbodies->add_constraint_point_with_angle_limit(anchor, b4, { -0.15, 0.0, 0.0 }, Eigen::Vector3d(0.0, 0.2, 1.0).normalized(), 2.0);
	b4.add_torque({ 3.0, 1.0, 0.5 });
	auto b5 = mk({ 0.0, 0.0, 0.4 }, 21.0);
	sim.rigidbodies->add_constraint_spring(anchor, b5, { 0.0, 0.01, 0.04 }, { 0.02, 0.0, 0.36 }, 250.0, 3.0);
	b5.add_force_at_centroid({ 2.0, 1.0, 6.0 });
	auto b6 = mk({ 0.3, 0.3, 0.0 }, -4.0);
	sim.rigidbodies->add_constraint_spring_with_limits(b2, b6, { 0.02, 0.3, 0.0 }, { 0.29, 0.31, 0.01 }, 120.0, 0.2, 0.2715, 1.0);
	b6.add_force_at_centroid({ 8.0, 0.0, -1.0 });
	return sc;
}

static Scene make_scene(const Args& a)
{
	if (a.scene == "zoo") return scene_zoo(a);
	if (a.scene == "joints") return scene_joints(a);
	if (a.scene == "tetchain") return scene_tetchain(a);
	if (a.scene == "boxes") return scene_boxes(a);
	if (a.scene == "attach") return scene_attach(a);
	if (a.scene == "tetdrop") return scene_tetdrop(a);
	if (a.scene == "magnet") return scene_magnet(a);
	if (a.scene == "tetbar") return scene_tetbar(a);
	if (a.scene == "cloth") return scene_cloth(a, false, 0.0);
	if (a.scene == "cloth_shells") return scene_cloth(a, true, 0.3);
	std::cerr << "unknown scene " << a.scene << "\n";
	exit(2);
}

// One time step exactly as Simulation::run does it (script cycle, then the step).
static void one_step(stark::Simulation& sim) { sim.run_one_time_step(); }

// ---------------------------------------------------------------------------------------------------
// dump of one Newton iteration on an injected state
// ---------------------------------------------------------------------------------------------------
static void dump_iteration(const Args& a, stark::Simulation& sim)
{
	stark::core::Stark& st = sim.get_stark();
	auto gp = st.global_potential;
	auto ctx = st.context;
	const int n_threads = ctx->n_threads;
	Dumper D(a.dump);
	D.str("scene", a.scene);
	D.integer("n", a.n);
	D.integer("steps_before_dump", a.steps);
	D.num("dt", st.dt);
	D.num("time", st.current_time);
	D.raw("gravity", "[" + std::to_string(st.gravity[0]) + "," + std::to_string(st.gravity[1]) + "," + std::to_string(st.gravity[2]) + "]");

	// Script + before_time_step exactly as the next step would do, then inject v1 := v0 as the DoF state
	// (the reference's own initial guess is v1 = 0; a non-zero state exercises every derivative term).
	sim.get_script().run_a_cycle(st.current_time);
	st.callbacks->run_before_time_step();
	{
		auto& dyn = *sim.deformables->point_sets;
		for (int i = 0; i < dyn.size(); i++) dyn.v1[i] = a.inject * dyn.v0[i];
		auto& rb = *sim.rigidbodies->rb;
		for (int i = 0; i < rb.get_n_bodies(); i++) { rb.v1[i] = a.inject * rb.v0[i]; rb.w1[i] = a.inject * rb.w0[i]; }
		// small deterministic perturbation so that rigid DoFs and symmetric nodes are not exactly zero
		for (int i = 0; i < dyn.size(); i++) {
			dyn.v1[i] += 1e-3 * Eigen::Vector3d(std::sin(0.37 * i), std::cos(0.91 * i), std::sin(1.3 * i + 0.5));
		}
		for (int i = 0; i < rb.get_n_bodies(); i++) {
			rb.v1[i] += 1e-4 * Eigen::Vector3d(0.3, -0.2, 0.5);
			rb.w1[i] += 1e-4 * Eigen::Vector3d(-0.1, 0.4, 0.2);
		}
	}
	st.callbacks->newton->run_before_energy_evaluation();

	// ---- DoFs
	const int ndofs = gp->get_total_n_dofs();
	D.integer("ndofs", ndofs);
	{
		std::ostringstream s; s << "[";
		const auto& offs = gp->get_dofs_offsets();
		for (size_t i = 0; i < offs.size(); i++) { if (i) s << ","; s << offs[i]; }
		s << "]";
		D.raw("dof_offsets", s.str());
		std::ostringstream ids; ids << "[";
		const auto& maps = gp->get_dof_maps();
		for (size_t i = 0; i < maps.size(); i++) { if (i) ids << ","; ids << "\"" << maps[i].id() << "\""; }
		ids << "]";
		D.raw("dof_array_ids", ids.str());
		std::vector<double> u(ndofs);
		gp->get_dofs(u.data());
		D.bin("dofs", u.data(), u.size());
	}

	// ---- the compiled potentials NewtonsMethod itself holds (a second instance would re-hash and re-JIT)
	symx::SecondOrderCompiledGlobal& compiled = *st.newton->compiled;

	// ---- potentials: inputs
	std::unordered_map<std::uintptr_t, int> array_index;
	std::ostringstream arrays_json; arrays_json << "[";
	std::ostringstream pots_json; pots_json << "[";
	const auto& pots = gp->get_potentials();
	bool first_array = true;
	for (size_t p = 0; p < pots.size(); p++) {
		const symx::Potential& pot = *pots[p];
		auto mws = pot.get_mws();
		const int n_elem = mws->conn.n_elements();
		const int stride = mws->conn.stride;
		if (p) pots_json << ",\n    ";
		pots_json << "{\"name\": \"" << pot.get_name() << "\", \"n_elements\": " << n_elem << ", \"conn_stride\": " << stride
			<< ", \"has_condition\": " << (pot.has_conditional() ? "true" : "false")
			<< ", \"n_symbols\": " << mws->ws.get_n_symbols();
		if (n_elem > 0) D.bin("pot" + std::to_string(p) + "_conn", mws->conn.data(), (size_t)n_elem * stride);
		pots_json << ", \"maps\": [";
		for (size_t m = 0; m < mws->maps.size(); m++) {
			const auto& map = mws->maps[m];
			const std::uintptr_t id = map.id();
			auto it = array_index.find(id);
			int aidx;
			if (it == array_index.end()) {
				aidx = (int)array_index.size();
				array_index[id] = aidx;
				const int n = map.n_elements();
				if (!first_array) arrays_json << ","; first_array = false;
				arrays_json << "{\"id\": \"" << id << "\", \"n_elements\": " << n << ", \"stride\": " << map.stride << "}";
				D.bin("array" + std::to_string(aidx), map.data(), (size_t)n * map.stride);
			} else {
				aidx = it->second;
			}
			if (m) pots_json << ", ";
			pots_json << "{\"array\": " << aidx << ", \"stride\": " << map.stride << ", \"conn_idx\": " << map.connectivity_index
				<< ", \"first_symbol\": " << map.first_symbol_idx << "}";
		}
		pots_json << "]";

		// DoF layout of the element (which conn column feeds each 3-block, and of which DoF set)
		auto& cp = *compiled.compiled_potentials[p];
		pots_json << ", \"n_dofs\": " << cp.n_dofs << ", \"dof_in_conn\": [";
		for (size_t b = 0; b < cp.dof_in_conn.size(); b++) {
			if (b) pots_json << ", ";
			pots_json << "[" << cp.dof_in_conn[b].dof_set << ", " << cp.dof_in_conn[b].conn_idx << "]";
		}
		pots_json << "]";

		// per-element outputs  [E | grad(n) | hess(n*n)]
		const int n_out = 1 + cp.n_dofs + cp.n_dofs * cp.n_dofs;
		if (n_elem > 0) {
			std::vector<double> sol((size_t)n_elem * n_out, 0.0);
			std::vector<uint8_t> active(n_elem, 1);
			if (pot.has_conditional()) {
				cp._evaluate_element_condition(n_threads);
				active = cp.has_element_positive_condition;
			}
			cp.P__dP_du__d2P_du2.run(n_threads,
				[&](const symx::View<double> s, int32_t e, int32_t tid, const symx::View<int32_t> conn) {
					for (int k = 0; k < n_out; k++) sol[(size_t)e * n_out + k] = s[k];
				}, cp.has_element_positive_condition);
			D.bin("pot" + std::to_string(p) + "_sol", sol.data(), sol.size());
			D.bin("pot" + std::to_string(p) + "_active", active.data(), active.size());
		}
		// user potentials: the operation sequences the reference's code generator prints (Compilation.cpp:381-469), as
		// SecondOrderCompiledPotential.cpp:9-80 builds them -- input of the GPU path's Sequence -> CUDA back-end
		if (pot.get_name() == a.ops) {
			std::vector<symx::Scalar> dofs;
			std::vector<int32_t> block_slots;
			for (const auto& dof_map : gp->get_dof_maps()) {
				const std::vector<symx::Scalar> set_dofs = mws->get_symbols(dof_map);
				for (size_t k = 0; k < set_dofs.size(); k += 3) block_slots.push_back(set_dofs[k].get_symbol_idx());
				dofs.insert(dofs.end(), set_dofs.begin(), set_dofs.end());
			}
			const symx::Scalar v = pot.get_expression();
			symx::DiffCache diff_cache;
			const symx::Vector g = symx::gradient(v, dofs, diff_cache);
			const symx::Matrix h = symx::gradient(g, dofs, /*symmetric=*/true, diff_cache);
			symx::Sequence seq_p({ v });
			symx::Sequence seq_pgh(symx::collect_scalars({ { v }, g.values(), h.values() }));
			auto dump_ops = [&](const symx::Sequence& seq, const std::string& tag) {
				std::vector<int32_t> ints; std::vector<double> consts;
				for (const auto& o : seq.ops) { ints.insert(ints.end(), { (int32_t)o.type, o.dst, o.a, o.b, o.cond }); consts.push_back(o.constant); }
				D.bin("pot" + std::to_string(p) + "_ops_" + tag, ints.data(), ints.size());
				D.bin("pot" + std::to_string(p) + "_opsc_" + tag, consts.data(), consts.size());
			};
			dump_ops(seq_p, "p"); dump_ops(seq_pgh, "pgh");
			D.bin("pot" + std::to_string(p) + "_block_slots", block_slots.data(), block_slots.size());
			pots_json << ", \"user_ops\": true, \"n_in\": " << seq_pgh.get_n_inputs();
		}
		pots_json << ", \"n_out\": " << n_out << "}";
	}
	arrays_json << "]";
	pots_json << "]";
	D.raw("arrays", arrays_json.str());
	D.raw("potentials", pots_json.str());

	// ---- global evaluation
	double E = 0.0, E_only = 0.0;
	Eigen::VectorXd grad(ndofs);
	compiled.evaluate_P(E_only);
	auto eh = compiled.evaluate_P__dP_du__local_d2P_du2(E, grad);
	D.num("E", E);
	D.num("E_only", E_only);
	D.num("residual_inf", grad.cwiseAbs().maxCoeff());
	D.bin("grad", grad.data(), (size_t)ndofs);
	D.integer("n_element_hessians", (long long)eh->size());

	// ---- assembly (unprojected, as PPN does first)
	auto hess = eh->assemble_global(n_threads, ndofs);
	{
		const auto& crs = *hess->crs_current;
		D.integer("bcsr_n_block_rows", hess->n_block_rows);
		D.integer("bcsr_nnzb", (long long)crs.cols.size());
		D.bin("bcsr_rows", crs.rows.data(), crs.rows.size());
		D.bin("bcsr_cols", crs.cols.data(), crs.cols.size());
		D.bin("bcsr_vals", crs.vals.data(), crs.vals.size());
	}

	// ---- linear solve (BDPCG, same forcing tolerance as NewtonsMethod::_solve_linear_system)
	{
		const double residual_norm = grad.cwiseAbs().maxCoeff();
		const double forcing = std::min(1e-2, residual_norm * std::min(0.5, std::sqrt(residual_norm)));
		const double abs_tol = std::max(forcing, st.settings.newton.cg_abs_tolerance);
		Eigen::VectorXd rhs = -grad, du = Eigen::VectorXd::Zero(ndofs);
		hess->set_preconditioner(bsm::Preconditioner::BlockDiagonal);
		hess->prepare_preconditioning(n_threads);
		bsm::PCGContext pc;
		bsm::PCGInfo info = bsm::solve_pcg(*hess, du.data(), rhs.data(), ndofs, abs_tol, st.settings.newton.cg_rel_tolerance,
			st.settings.newton.cg_max_iterations, n_threads, st.settings.newton.cg_stop_on_indefiniteness, pc);
		D.num("pcg_abs_tol", abs_tol);
		D.num("pcg_rel_tol", st.settings.newton.cg_rel_tolerance);
		D.integer("pcg_iterations", info.n_iterations);
		D.integer("pcg_converged", info.converged ? 1 : 0);
		D.num("du_dot_grad", du.dot(grad));
		D.bin("pcg_du", du.data(), (size_t)ndofs);
	}

	// ---- PD projection of every element Hessian (clamp, eps from settings)
	{
		const double eps = st.settings.newton.projection_eps;
		std::vector<double> proj;
		std::vector<int32_t> sizes;
		for (size_t i = 0; i < eh->hessians.size(); i++) {
			auto& h = eh->hessians[i];
			const int n = 3 * h.n_blocks_per_dim;
			std::vector<double> m(h.values, h.values + n * n);
			project_to_PD_inplace(m.data(), n, eps, false);
			proj.insert(proj.end(), m.begin(), m.end());
			sizes.push_back(n);
		}
		D.bin("projected_hessians", proj.data(), proj.size());
		D.bin("projected_sizes", sizes.data(), sizes.size());
		// element order of ElementHessians (block rows) so the fixtures can be matched up
		std::vector<int32_t> rows;
		for (size_t i = 0; i < eh->hessians.size(); i++) {
			auto& h = eh->hessians[i];
			for (int b = 0; b < h.n_blocks_per_dim; b++) rows.push_back(h.block_rows[b]);
		}
		D.bin("element_block_rows", rows.data(), rows.size());
	}

	// ---- collision detection
	auto contact = sim.interactions->contact;
	if (contact->is_initialized && !contact->is_empty()) {
		std::ostringstream mj; mj << "[";
		for (size_t g = 0; g < contact->meshes.size(); g++) {
			auto& mesh = contact->meshes[g];
			if (g) mj << ", ";
			const bool dummy_tri = (mesh.loc_triangles.size() == 1 && mesh.loc_triangles[0][0] == -1);
			mj << "{\"ps\": " << (mesh.ps == stark::PhysicalSystem::Deformable ? 0 : 1) << ", \"idx_in_ps\": " << mesh.idx_in_ps
				<< ", \"n_vertices\": " << mesh.vertices.size() << ", \"n_triangles\": " << (dummy_tri ? 0 : mesh.loc_triangles.size())
				<< ", \"n_edges\": " << mesh.loc_edges.size() << ", \"contact_thickness\": " << std::setprecision(17) << contact->contact_thicknesses[g] << "}";
			D.bin("mesh" + std::to_string(g) + "_vertices", &mesh.vertices[0][0], mesh.vertices.size() * 3);
			if (!dummy_tri) D.bin("mesh" + std::to_string(g) + "_triangles", &mesh.loc_triangles[0][0], mesh.loc_triangles.size() * 3);
			D.bin("mesh" + std::to_string(g) + "_edges", &mesh.loc_edges[0][0], mesh.loc_edges.size() * 2);
			// index of every collision vertex in its physical system (deformable: global node, rigid: row of rigidbody_local_vertices)
			std::vector<int32_t> ps_index(mesh.vertices.size());
			for (int i = 0; i < (int)mesh.vertices.size(); i++) ps_index[i] = contact->_local_to_ps_global_indices<1>((int)g, { i })[0];
			D.bin("mesh" + std::to_string(g) + "_ps_index", ps_index.data(), ps_index.size());
		}
		D.bin("rigidbody_local_vertices", &contact->rigidbody_local_vertices.data[0][0], contact->rigidbody_local_vertices.data.size() * 3);
		{
			std::ostringstream fj; fj << "[";
			bool ff = true;
			for (const auto& kv : contact->pair_coulombs_mu) { if (!ff) fj << ", "; ff = false; fj << "[" << kv.first[0] << ", " << kv.first[1] << ", " << std::setprecision(17) << kv.second << "]"; }
			fj << "]";
			D.raw("friction_pairs", fj.str());
			D.num("friction_stick_slide_threshold", contact->global_params.friction_stick_slide_threshold);
		}
		mj << "]";
		D.raw("meshes", mj.str());
		std::ostringstream bl; bl << "[";
		bool fb = true;
		for (const auto& pr : contact->disabled_collision_pairs) { if (!fb) bl << ", "; fb = false; bl << "[" << pr[0] << ", " << pr[1] << "]"; }
		bl << "]";
		D.raw("blacklist", bl.str());
		D.num("contact_stiffness", contact->contact_stiffness);

		const double max_thickness = *std::max_element(contact->contact_thicknesses.begin(), contact->contact_thicknesses.end());
		const double enl = 2.0 * max_thickness;
		D.num("proximity_enlargement", enl);
		const tmcd::ProximityResults& pr = contact->_run_proximity_detection(st, st.dt);
		auto dump_list = [&](const std::string& name, const std::vector<std::vector<int32_t>>& rows, const std::vector<double>& dist) {
			std::vector<int32_t> flat;
			for (auto& r : rows) flat.insert(flat.end(), r.begin(), r.end());
			D.bin("prox_" + name + "_ids", flat.data(), flat.size());
			D.bin("prox_" + name + "_dist", dist.data(), dist.size());
			D.integer("prox_" + name + "_n", (long long)rows.size());
			D.integer("prox_" + name + "_width", rows.empty() ? 0 : (long long)rows[0].size());
		};
		{
			std::vector<std::vector<int32_t>> r; std::vector<double> d;
			for (auto& x : pr.point_triangle.point_point) { r.push_back({ x.first.set, x.first.idx, x.second.triangle.set, x.second.triangle.idx, x.second.point.idx }); d.push_back(x.distance); }
			dump_list("pt_pp", r, d);
		}
		{
			std::vector<std::vector<int32_t>> r; std::vector<double> d;
			for (auto& x : pr.point_triangle.point_edge) { r.push_back({ x.first.set, x.first.idx, x.second.triangle.set, x.second.triangle.idx, x.second.edge.vertices[0], x.second.edge.vertices[1] }); d.push_back(x.distance); }
			dump_list("pt_pe", r, d);
		}
		{
			std::vector<std::vector<int32_t>> r; std::vector<double> d;
			for (auto& x : pr.point_triangle.point_triangle) { r.push_back({ x.first.set, x.first.idx, x.second.set, x.second.idx }); d.push_back(x.distance); }
			dump_list("pt_pt", r, d);
		}
		{
			std::vector<std::vector<int32_t>> r; std::vector<double> d;
			for (auto& x : pr.edge_edge.point_point) { r.push_back({ x.first.edge.set, x.first.edge.idx, x.first.point.idx, x.second.edge.set, x.second.edge.idx, x.second.point.idx }); d.push_back(x.distance); }
			dump_list("ee_pp", r, d);
		}
		{
			std::vector<std::vector<int32_t>> r; std::vector<double> d;
			for (auto& x : pr.edge_edge.point_edge) { r.push_back({ x.first.edge.set, x.first.edge.idx, x.first.point.idx, x.second.set, x.second.idx }); d.push_back(x.distance); }
			dump_list("ee_pe", r, d);
		}
		{
			std::vector<std::vector<int32_t>> r; std::vector<double> d;
			for (auto& x : pr.edge_edge.edge_edge) { r.push_back({ x.first.set, x.first.idx, x.second.set, x.second.idx }); d.push_back(x.distance); }
			dump_list("ee_ee", r, d);
		}
		const tmcd::IntersectionResults& ir = contact->_run_intersection_detection(st, st.dt);
		{
			std::vector<int32_t> flat;
			for (auto& x : ir.edge_triangle) { flat.push_back(x.first.set); flat.push_back(x.first.idx); flat.push_back(x.second.set); flat.push_back(x.second.idx); }
			D.bin("intersections", flat.data(), flat.size());
			D.integer("n_intersections", (long long)ir.edge_triangle.size());
		}
	}
	D.finish();
}

// ---------------------------------------------------------------------------------------------------
int main(int argc, char** argv)
{
	Args a = parse(argc, argv);
	Scene sc = make_scene(a);
	stark::Simulation& sim = *sc.sim;
	stark::core::Stark& st = sim.get_stark();

	if (!a.dump.empty()) {
		for (int s = 0; s < a.steps; s++) one_step(sim);
		dump_iteration(a, sim);

		// residual trace of the next real step (for end-to-end comparison)
		std::vector<double> residuals;
		st.callbacks->newton->set_residual([&](Eigen::VectorXd& r) { const double v = r.cwiseAbs().maxCoeff(); residuals.push_back(v); return v; });
		// restore v1 := 0 happens inside before_time_step of the next step
		one_step(sim);
		std::ofstream f(a.dump + "/next_step_residuals.txt");
		f << std::setprecision(17);
		for (double r : residuals) f << r << "\n";
		auto stats = st.newton->get_last_solve_stats();
		std::ofstream g(a.dump + "/next_step_stats.json");
		g << "{\"newton_iterations\": " << stats.newton_iterations << ", \"cg_iterations\": " << stats.cg_iterations
			<< ", \"ls_inv\": " << stats.ls_inv_iterations << ", \"ls_bt\": " << stats.ls_bt_iterations << "}\n";
		return 0;
	}

	if (!a.trace.empty()) {
		// trajectory trace: the same scene through two builds (the unmodified reference / the reference with the stark_b200
		// NewtonsMethod linked in, integration/Makefile.shim) must agree step by step
		std::ofstream f(a.trace);
		f << std::setprecision(17);
		for (int s = 0; s < a.steps; s++) {
			const int step_before = st.current_time_step;
			one_step(sim);
			auto stats = st.newton->get_last_solve_stats();
			auto& dyn = *sim.deformables->point_sets;
			auto& rb = *sim.rigidbodies->rb;
			double sx = 0.0, sx2 = 0.0, mx = 0.0;
			for (int i = 0; i < dyn.size(); i++) for (int c = 0; c < 3; c++) { const double x = dyn.x0[i][c]; sx += x; sx2 += x * x; mx = std::max(mx, std::abs(x)); }
			double rt = 0.0, rq = 0.0;
			for (int i = 0; i < rb.get_n_bodies(); i++) { for (int c = 0; c < 3; c++) rt += (i + 1) * (c + 1) * rb.t0[i][c]; rq += (i + 1) * (rb.q0[i].w() + 2.0 * rb.q0[i].x() + 3.0 * rb.q0[i].y() + 4.0 * rb.q0[i].z()); }
			f << "{\"step\": " << s << ", \"time\": " << st.current_time << ", \"accepted\": " << ((st.current_time_step > step_before) ? 1 : 0)
				<< ", \"newton_iterations\": " << stats.newton_iterations << ", \"cg_iterations\": " << stats.cg_iterations
				<< ", \"ls_inv\": " << stats.ls_inv_iterations << ", \"ls_bt\": " << stats.ls_bt_iterations
				<< ", \"n_points\": " << dyn.size() << ", \"sum_x\": " << sx << ", \"sum_x2\": " << sx2 << ", \"max_abs_x\": " << mx
				<< ", \"rigid_t\": " << rt << ", \"rigid_q\": " << rq << "}" << std::endl;
		}
		return 0;
	}

	if (a.bench) {
		// warm-up (includes JIT/initialization)
		for (int s = 0; s < a.warmup; s++) one_step(sim);
		auto& logger = sim.get_logger();
		long long iters = 0, cg = 0, evals = 0;
		int n_evals = 0;
		st.callbacks->newton->set_residual([&](Eigen::VectorXd& r) { n_evals++; return r.cwiseAbs().maxCoeff(); });
		const double t0 = omp_get_wtime();
		double solve_time = 0.0;
		int accepted = 0;
		for (int s = 0; s < a.steps; s++) {
			const double ta = omp_get_wtime();
			const int step_before = st.current_time_step;
			one_step(sim);
			solve_time += omp_get_wtime() - ta;
			auto stats = st.newton->get_last_solve_stats();
			iters += stats.newton_iterations;
			cg += stats.cg_iterations;
			accepted += (st.current_time_step > step_before) ? 1 : 0;
		}
		const double wall = omp_get_wtime() - t0;
		std::cout << std::setprecision(10)
			<< "{\"impl\": \"reference\", \"scene\": \"" << a.scene << "\", \"n\": " << a.n
			<< ", \"ndofs\": " << st.global_potential->get_total_n_dofs()
			<< ", \"threads\": " << st.context->n_threads
			<< ", \"steps\": " << a.steps << ", \"accepted_steps\": " << accepted
			<< ", \"newton_iterations\": " << iters << ", \"evaluations\": " << n_evals << ", \"cg_iterations\": " << cg
			<< ", \"wall_s\": " << wall
			<< ", \"newton_it_per_s\": " << (wall > 0 ? iters / wall : 0.0)
			<< ", \"time\": " << st.current_time << "}" << std::endl;
		return 0;
	}

	std::cerr << "nothing to do: pass --dump DIR or --bench\n";
	return 2;
}
