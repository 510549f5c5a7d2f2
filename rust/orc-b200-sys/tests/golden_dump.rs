//! UNBUILT / UNTESTED here (no cargo / rustc in the build image). What it is for: pinning the CPU oracle of the orc_b200 repo
//! (oracle/, a C++ restatement of ORC's steady SIMPLE path) to the REAL reference, bit for bit.
//!
//!   cd rust/orc-b200-sys && cargo test --release --test golden_dump -- --nocapture
//!
//! runs the unmodified `orc::solver::solve_steady` on the two example meshes of the reference with exactly the boundary conditions,
//! settings and iteration counts of tests/golden/make_golden.py (which produced tests/golden/kat_<name>.npz from the oracle), and
//! writes u, v, w, p as little-endian f64 to tests/golden/orc_dump_<name>.bin (header: u64 cell count, u64 iterations).
//! `python -m pytest tests/test_oracle_kats.py -k real_orc_dump` then compares the oracle's fields with the dump (`array_equal`).
//! With `--features gpu` (liborc_b200.so built, a B200 present) the same cases also run through the shim's `solve_steady` in
//! reference-order mode and must reproduce the dump bit for bit.
use nalgebra::DVector;
use orc::io::read_mesh;
use orc::mesh::{FaceConditionTypes, Mesh};
use orc::numerical_types::{Float, Uint, Vector};
use orc::settings::*;
use std::io::Write;

struct Case {
    name: &'static str,
    walls: &'static [&'static str],
    moving: Option<&'static str>,
    dp_dx: Float,
    u_wall: Float,
    tvd_quick: bool,
    iterations: Uint,
}

const CASES: [Case; 2] = [
    Case { name: "channel_flow", walls: &["WALL"], moving: None, dp_dx: 5.0, u_wall: 0.0, tvd_quick: true, iterations: 3 },
    Case { name: "couette_flow_128x64x1", walls: &["TOP_WALL", "BOTTOM_WALL"], moving: Some("TOP_WALL"), dp_dx: 10.0, u_wall: 5e-4, tvd_quick: false, iterations: 2 },
];

/// tests/cases.py: couette_bcs (the boundary conditions of src/tests.rs:60-76)
fn set_bcs(mesh: &mut Mesh, c: &Case) {
    for z in c.walls {
        mesh.get_face_zone(z).zone_type = FaceConditionTypes::Wall;
        mesh.get_face_zone(z).vector_value = Vector { x: if Some(*z) == c.moving { c.u_wall } else { 0. }, y: 0., z: 0. };
    }
    mesh.get_face_zone("INLET").zone_type = FaceConditionTypes::PressureInlet;
    mesh.get_face_zone("INLET").scalar_value = -c.dp_dx * 0.002;
    mesh.get_face_zone("OUTLET").zone_type = FaceConditionTypes::PressureOutlet;
    mesh.get_face_zone("OUTLET").scalar_value = 0.;
    mesh.get_face_zone("PERIODIC_-Z").zone_type = FaceConditionTypes::Symmetry;
    mesh.get_face_zone("PERIODIC_+Z").zone_type = FaceConditionTypes::Symmetry;
}

fn settings(c: &Case) -> NumericalSettings {
    let mut s = NumericalSettings::default();
    if c.tvd_quick {
        s.momentum = MomentumDiscretization::TVD(TVD_QUICK);
    }
    s
}

fn reference_root() -> String {
    std::env::var("ORC_REFERENCE_ROOT").unwrap_or_else(|_| format!("{}/../../../reference", env!("CARGO_MANIFEST_DIR")))
}

fn dump(path: &str, iterations: Uint, fields: [&DVector<Float>; 4]) {
    let mut f = std::fs::File::create(path).expect("cannot create the dump");
    f.write_all(&(fields[0].len() as u64).to_le_bytes()).unwrap();
    f.write_all(&(iterations as u64).to_le_bytes()).unwrap();
    for v in fields {
        for x in v.iter() {
            f.write_all(&x.to_le_bytes()).unwrap();
        }
    }
}

#[test]
fn golden_dump_of_the_real_reference() {
    let out_dir = format!("{}/../../tests/golden", env!("CARGO_MANIFEST_DIR"));
    for c in CASES.iter() {
        let mut mesh = read_mesh(&format!("{}/examples/{}.msh", reference_root(), c.name));
        set_bcs(&mut mesh, c);
        let n = mesh.cells.len();
        let (mut u, mut v, mut w, mut p) = (DVector::zeros(n), DVector::zeros(n), DVector::zeros(n), DVector::zeros(n));
        orc::solver::solve_steady(&mut mesh, &mut u, &mut v, &mut w, &mut p, &settings(c), 1000.0, 1e-3, c.iterations, 1);
        dump(&format!("{out_dir}/orc_dump_{}.bin", c.name), c.iterations, [&u, &v, &w, &p]);
        println!("{}: {} cells, {} iterations, mean u = {:e}", c.name, n, c.iterations, u.mean());

        #[cfg(feature = "gpu")]
        {
            // the B200 path behind the reference's own signature; ORC_REDUCE_AUTO picks reference-order reductions on these meshes
            let mut mesh2 = read_mesh(&format!("{}/examples/{}.msh", reference_root(), c.name));
            set_bcs(&mut mesh2, c);
            let (mut u2, mut v2, mut w2, mut p2) = (DVector::zeros(n), DVector::zeros(n), DVector::zeros(n), DVector::zeros(n));
            orc_b200_sys::solve_steady(&mut mesh2, &mut u2, &mut v2, &mut w2, &mut p2, &settings(c), 1000.0, 1e-3, c.iterations, 1);
            for (a, b) in [(&u, &u2), (&v, &v2), (&w, &w2), (&p, &p2)] {
                assert!(a.iter().zip(b.iter()).all(|(x, y)| x.to_bits() == y.to_bits()), "{}: GPU path differs from ORC", c.name);
            }
        }
    }
}
