// Replacement for the body of the render thread in render() (src/render/mod.rs:984-1026).  Signature, the cancel-watcher
// thread (:947-958), the progress thread (:965-982), the PPM writer (:1031-1088) and RenderDone stay exactly as they are.
// UNTESTED IN THIS REPO (no cargo here).
//
//     let render_thread_handle = s.spawn(move || {
//         let scene = &render_config.scene;
//         println!("Rendering scene {} ({} objects), {} samples per pixel, {}x{} resolution (B200 backend)", ...);
//
            let (objs, tris, cam) = ffi::flatten(scene);
            let desc = ffi::ptb_scene_desc { objects: objs.as_ptr(), n_objects: objs.len() as u64,
                                             triangles: tris.as_ptr(), n_triangles: tris.len() as u64, camera: cam };
            let mut ctx: *mut ffi::ptb_ctx = std::ptr::null_mut();
            unsafe {
                assert_eq!(ffi::ptb_create(0, &mut ctx), ffi::PTB_OK, "no B200: the backend has no CPU fallback");
                assert_eq!(ffi::ptb_upload_scene(ctx, &desc), ffi::PTB_OK);
            }
            // the two helper threads keep working: stop_render is polled through `cancel`, progress through `samples_done`
            let cancel = std::sync::atomic::AtomicI32::new(0);
            let samples_done = std::sync::atomic::AtomicU64::new(0);
            let mut out = vec![0f32; grid_size * 3];
            let spp = render_config.samples_per_pixel as u64;
            let seed: u64 = rand::random();                       // the reference is OS-seeded too (mod.rs:53)
            // a small watcher copies stop_render -> cancel and samples_done -> processed_pixel_count every 100 ms (omitted)
            let rc = unsafe {
                ffi::ptb_render(ctx, res.width as i32, res.height as i32, 0, spp, seed, ffi::PTB_OUT_MEAN, out.as_mut_ptr(),
                                cancel.as_ptr() as *const i32, samples_done.as_ptr())
            };
            assert!(rc >= 0, "ptb_render failed");
            {   // same buffer order as the rayon path: index i <-> x = i % W, y = H-1 - i / W (mod.rs:805-806)
                let mut px = pixels.lock().unwrap();
                for (i, p) in px.iter_mut().enumerate() { *p = Vec3::new(out[3 * i], out[3 * i + 1], out[3 * i + 2]); }
            }
            unsafe { ffi::ptb_destroy(ctx) };
//         println!("Rendering complete");
//         stop_render.store(true, atomic::Ordering::Relaxed);
//         ... PPM writer and Image::new unchanged ...
